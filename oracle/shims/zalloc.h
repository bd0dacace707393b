/* TEST INFRASTRUCTURE (oracle build shim) -- not product code.
 *
 * Force-included (gcc -include) when building the *deterministic* variant of the
 * reference's c/mem.c.  The reference applies Numerical-Recipes 1-based indexing to
 * 0-based malloc buffers and reads Data[N], one element past the end
 * (/root/reference/c/mem.c:210,214).  Replacing malloc with a zero-filling, slightly
 * over-sized calloc makes that element read as 0.0, which is the canonical semantics
 * this repository tests against.
 * The reference source itself is compiled unmodified.
 */
#include <stdlib.h>
#define malloc(n) calloc(1, (n) + 16)
