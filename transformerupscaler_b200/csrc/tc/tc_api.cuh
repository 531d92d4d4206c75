// Internal interface between the C-ABI layer and the tcgen05 (5th-gen tensor core) kernels.
#pragma once
#include "../tu_common.cuh"

namespace tu {

constexpr int TU_TC_UNSUPPORTED = 1;   // shape not covered by the tensor-core kernel: caller falls back to SIMT

// 1 when the tcgen05 kernels were compiled in and may be used on this device.
int tc_available();
// debug: 0 = UMMA descriptor base_offset 0, 1 = base_offset (addr>>7)&7 for row-shifted operand views
void tc_set_base_off_mode(int m);

// 3x3 / pad 1 convolution, 64 input channels, NHWC bf16, fp32 accumulation in TMEM.
// w: [chunk][tap][co 64][ci 64] bf16.  Returns TU_OK, an error code, or TU_TC_UNSUPPORTED.
int tc_conv3x3_c64(const bf16 *in, const bf16 *w, const float *bias, bf16 *out, int B, int H, int W, int stride, int relu,
                   int nchunk, int ps_r, cudaStream_t st);

}  // namespace tu
