/* chiron_b200 -- TEST-ONLY hooks.  Host-compiled instantiations of the sequential __host__ __device__ routines the GPU
 * kernels run (chiron_b200/csrc/cb_seq_algos.cuh), so that their logic can be unit-tested on a machine without a GPU.
 * Nothing in the product path (chiron_b200/*.py, the cb_* entry points of chiron_b200.h) calls these. */
#ifndef CHIRON_B200_SELFTEST_H
#define CHIRON_B200_SELFTEST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* The beam-search routine of beam_kernel on one window: logits[len][n_class] -> out[], returns the decoded length
 * (-2: node pool exhausted, -1: bad arguments).  `pool` bounds the trie node pool (exercises in-place compaction). */
int cb_selftest_beam(const float* logits, int len, int n_class, int beam_width, int pool, int8_t* out);
/* The same routine over the 16-bit trie nodes of the shared-memory kernels (pool, len, beam_width < 32768). */
int cb_selftest_beam16(const float* logits, int len, int n_class, int beam_width, int pool, int8_t* out);

/* The displacement routine of asm_disp_kernel on one adjacent window pair (kernel = CB_ASM_*). */
int cb_selftest_disp(const int8_t* cur, int la, const int8_t* prev, int lb, int kernel, int jump, int L);

#ifdef __cplusplus
}
#endif
#endif
