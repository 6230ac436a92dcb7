/* TEST INFRASTRUCTURE (oracle build only). */
#ifndef ODX_STUB_GSL_SF_H
#define ODX_STUB_GSL_SF_H
double gsl_sf_choose(unsigned int n, unsigned int m);
#endif
