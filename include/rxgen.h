/* rxgen.h -- C-ABI of librxgen.so: benchmark INPUT generators, host-only, no CUDA.
 *
 * Kept out of librxcuda.so on purpose: the generators are not part of the accelerated path (SURVEY.md
 * section 2a marks them out of scope as a product), and bench.py's reference arm must be able to build
 * its input graph without mapping the product library.
 */
#ifndef RXGEN_H
#define RXGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RXG_ERR_INVALID (-1)

/* ---- benchmark inputs (host-side restatement of the reference's seeded generators;
 *      rustworkx-core/src/generators/random_graph.rs:249-345,402-466,756-824).  Each writes at most
 *      `cap` edges and returns the number of edges generated (or a negative value: -1 invalid argument / `cap` too small). -------------- */
int64_t rxc_gen_gnp(uint32_t n, double p, uint64_t seed, int directed, uint32_t *src,
                    uint32_t *dst, uint64_t cap);
int64_t rxc_gen_gnm(uint32_t n, uint64_t m, uint64_t seed, int directed, uint32_t *src,
                    uint32_t *dst, uint64_t cap);
int64_t rxc_gen_barabasi_albert(uint32_t n, uint32_t m, uint64_t seed, uint32_t *src,
                                uint32_t *dst, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* RXGEN_H */
