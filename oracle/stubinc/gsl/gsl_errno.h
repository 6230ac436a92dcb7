/* TEST INFRASTRUCTURE (oracle build only). Declaration-only stand-in for <gsl/gsl_errno.h>:
 * GSL is not installed in this image; the unchanged reference sources only need these names. */
#ifndef ODX_STUB_GSL_ERRNO_H
#define ODX_STUB_GSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_ERANGE = 2,
       GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5, GSL_ENOMEM = 8, GSL_EBADFUNC = 9,
       GSL_ENOPROG = 27, GSL_ENOPROGJ = 28 };
const char *gsl_strerror(const int gsl_errno);
#endif
