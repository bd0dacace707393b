/* TEST INFRASTRUCTURE (oracle build shim) -- not product code.
 *
 * Force-included (gcc -include) when building the *intended-semantics* variant of the
 * reference's c/correlation.c.  The shipped code calls cexp(_Cbuild(w*tau, 1)), i.e.
 * exp(w*tau + 1i) (/root/reference/c/correlation.c:158,162,163,170); the intended
 * weight, visible in the comments at c/correlation.c:166-167,172, is exp(i*w*tau).
 * This shim maps cexp(z) -> cexp(i*Re z) so the unmodified source evaluates the
 * intended formula.  Used only to pin the weight="intended" mode of the CUDA kernel.
 */
#include <complex.h>
static inline double _Complex dp_fix_cexp(double _Complex z) { return cexp(creal(z) * 1.0j); }
#define cexp(z) dp_fix_cexp(z)
