/* Bring-up / measurement probes of the tcgen05 + TMA path (tools/umma_probe.py, tools/umma_rate.py,
 * tools/umma_ts.py).  A TOOLS library (tools/probe/libseldq_probe.so, built by tools/probe/build.sh): nothing of it
 * is linked into the product library libseldq.so or declared in include/seldq.h. */
#ifndef SELDQ_PROBE_H_
#define SELDQ_PROBE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
const char* seldq_probe_last_error(void);
int seldq_probe_tensor_map(void* host_map_128B, const void* gaddr, int32_t elem_bytes, int32_t rank,
                           const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                           int32_t swizzle);
int seldq_probe_tma_load(const void* host_map_128B, const void* dev_map_128B, int32_t rank,
                         const int32_t* coords, uint32_t box_bytes, uint32_t smem_offset,
                         void* out_smem_dump, uint32_t dump_bytes, void* stream);
int seldq_probe_umma(const void* a_image, uint32_t a_bytes, const void* b_image, uint32_t b_bytes,
                     uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int32_t n_mma,
                     uint32_t a_desc_step, uint32_t b_desc_step, int32_t n_cols,
                     float* out_128xN, void* stream);
/* tcgen05.mma issue-rate probes (tools/umma_rate.py, tools/umma_ts.py): cycles per MMA versus N for the product
 * kernels' operand layouts, and the A-from-tensor-memory variant (tcgen05.cp + TS-form MMA).  out: 2 x blocks int64
 * (issue cycles, issue + drain cycles); mode 0 of the second probe returns mismatch / non-zero counts instead. */
int seldq_probe_umma_rate(uint32_t n, int32_t n_mma, int32_t d_cycle, int32_t mode, int32_t blocks, void* out,
                          void* stream);
int seldq_probe_umma_ts(uint32_t n, int32_t n_slabs, int32_t g, int32_t mode, int32_t nbuf, int32_t blocks, void* out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELDQ_PROBE_H_ */
