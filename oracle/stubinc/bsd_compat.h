/* TEST INFRASTRUCTURE (oracle build only). Prototypes of the two BSD libc functions the macOS-targeted reference calls; without a
 * prototype mergesort(tstops, nbr_stops, ...) with nbr_stops declared double would pass its count in a floating-point register. */
#ifndef ODX_STUB_BSD_COMPAT_H
#define ODX_STUB_BSD_COMPAT_H
#include <stddef.h>
int mergesort(void *base, size_t nel, size_t width, int (*compar)(const void *, const void *));
int isnumber(int c);
#endif
