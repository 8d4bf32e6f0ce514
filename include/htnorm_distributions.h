/* htnorm-b200: compatibility shim for callers that include the reference's PRIVATE header
 * src/htnorm_distributions.h for the one public thing it holds: the allocation-failure code.
 *
 *   HTNORM_ALLOC_ERROR  <-> reference src/htnorm_distributions.h:19, read by the Cython wrapper
 *                           (pyhtnorm/_htnorm.pyx:47-48: `cdef extern from "htnorm_distributions.h"`)
 *
 * The rest of that header (mvn_output_t, mvn_rand_cov, mvn_rand_prec: src/htnorm_distributions.h:24-107) is the
 * reference's internal scratch / CPU sampling layer; here those steps are GPU kernels behind
 * htn_hyperplane_truncated_mvn / htn_structured_precision_mvn and have no host-callable counterpart.
 * With this file on the include path, pyhtnorm/_htnorm.pyx compiles UNCHANGED against libhtnorm_b200.so
 * (oracle/build_pyhtnorm.sh, tests/test_reference_wrapper.py).
 */
#ifndef HTNORM_DISTRIBUTIONS_H
#define HTNORM_DISTRIBUTIONS_H

#ifndef HTNORM_ALLOC_ERROR
#define HTNORM_ALLOC_ERROR -100
#endif

#endif
