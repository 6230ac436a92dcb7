/* odx_events.cuh — stochastic birth / death on the resident population (SSA single event, tau-leaping). */
#pragma once
extern "C" {
int odx_ssa_apply(odx_handle *h, double t, double u, odx_event_row *row) {
  (void)t; (void)u; (void)row; h->err = "odx_ssa_apply: not built yet"; return ODX_EUNSUP;
}
int odx_tauleap_apply(odx_handle *h, double t, double dt, uint64_t leap_index, odx_event_row *rows, int64_t max_rows, int64_t *nrows) {
  (void)t; (void)dt; (void)leap_index; (void)rows; (void)max_rows; (void)nrows; h->err = "odx_tauleap_apply: not built yet"; return ODX_EUNSUP;
}
}
