#!/usr/bin/env python3
"""Add the B200 branch to a COPY of the reference's src/methods_odexp.c (the reference tree itself is never modified and no
reference source is kept in this repository: the copy is written where the caller says, normally a git-ignored build directory).

    python integration/apply_gpu_branch.py <reference>/src/methods_odexp.c <out>/methods_odexp_gpu.c [--diff <out>/odesolver_gpu.patch]

Two insertions, both found by anchors that are stable parts of odesolver() (src/methods_odexp.c:83-763):
  1. after the #include block: the prototype of odx_gpu_odesolver() (integration/odexp_gpu_branch.c);
  2. in front of the `/* Initialize SIM */` block (:250), i.e. after the step-method name and the stopping times have been parsed
     and before any output file is opened: hand the run to the GPU branch when it accepts it.
With --diff the same change is also written as a unified diff for a maintainer to apply with patch(1)."""
import difflib, re, sys

PROTO = '''
/* B200 branch: integration/odexp_gpu_branch.c (returns 1 when it ran the simulation on the GPU) */
int odx_gpu_odesolver(int solver, double_array *tspan, int *status_out);
'''
CALL = '''    /* B200 branch: population mode with a device model registered in ODEXP_CUDA_MODEL (integration/odexp_gpu_branch.c) */
    if ( odx_gpu_odesolver((int)solver, tspan, &status) )
    {
        free(tstops);
        return status;
    }

'''


def apply(text):
    m = re.search(r'#include "rand_gen.h"[^\n]*\n', text)
    if not m:
        raise SystemExit("anchor 1 (#include \"rand_gen.h\") not found: not the methods_odexp.c this script knows")
    text = text[:m.end()] + PROTO + text[m.end():]
    m = re.search(r'^[ \t]*/\* Initialize SIM \*/[ \t]*\n', text, flags=re.M)
    if not m:
        raise SystemExit("anchor 2 (/* Initialize SIM */) not found")
    # `solver` is left unset by the reference when the method name is unknown (it falls back to rkck): give it that value
    text = text[:m.start()] + CALL + text[m.start():]
    text = re.sub(r'(\n[ \t]*enum solver solver)(;)', r'\1 = GSL_RKCK\2', text, count=1)
    return text


def main():
    src, dst = sys.argv[1], sys.argv[2]
    old = open(src, encoding="utf8", errors="surrogateescape").read()
    new = apply(old)
    open(dst, "w", encoding="utf8", errors="surrogateescape").write(new)
    if "--diff" in sys.argv:
        out = sys.argv[sys.argv.index("--diff") + 1]
        d = difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/src/methods_odexp.c", "b/src/methods_odexp.c", n=1)
        open(out, "w", encoding="utf8", errors="surrogateescape").writelines(d)


if __name__ == "__main__":
    main()
