/* xview_b200_measure - measurement and test hooks of libxview_b200.so.
 *
 * NOT part of the drop-in boundary (include/xview_b200.h): nothing here replaces a reference
 * statement.  bench.py, tools/ and the parity tests use these entry points to select kernel
 * variants, count launches and time the tensor-core convolutions on their own stream.
 */
#ifndef XVIEW_B200_MEASURE_H_
#define XVIEW_B200_MEASURE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- measurement hooks (no reference counterpart; used by bench.py) ------------------ */
/* Debug switches for parity tests: bit0 = never fuse max pooling into the conv epilogue,
 * bit1 = never use the transposed-role conv kernel, bit2 = conv1_1 through a materialised
 * operand buffer instead of in-kernel packing, bit3 = weight gradients on the CUDA cores instead
 * of the tensor-core kernel, bit5 = conv1_1
 * with global loads in the operand packers instead of TMA-staged input patches, bit6 = 3x3
 * convolutions load nine shifted tiles per channel chunk instead of three patch copies,
 * bit7 = single-CTA MMAs for the Cout >= 256 layers instead of the CTA-pair (cta_group::2)
 * kernel, bit8 = scalar (4 elements per thread) dropout kernel, bit9 = MC decode with two
 * barriers per sample instead of the eight-sample staging, bit10 = conv1_2 + pool1 on the
 * transposed-role kernel instead of the row-pair kernel, bit11 = single-CTA weight-gradient
 * kernel for every layer, bit12 = fit() loss head as decode + ce_grad + upsample8_transpose.
 * 0 = production behaviour. */
int xv_set_debug_flags(int flags);
/* Number of kernels this library has launched since it was loaded. */
int xv_launch_count(int64_t* out);
/* on != 0: bracket every tensor-core convolution launch with CUDA events on its stream;
 * on == 0: stop.  Either call discards the samples collected so far. */
int xv_profile_enable(int on);
/* Sum over the collected samples: device milliseconds, algorithmic FLOPs (2*MACs), launches. */
int xv_profile_read(double* ms_out, double* flops_out, int64_t* launches_out);

/* Times `iters` launches of one tensor-core convolution layer on synthetic bf16 data
 * (kernel study only; `flags` selects timing experiments, 0 = the production kernel). */
int xv_bench_conv_igemm(int n, int h, int w, int cin, int cout, int k, int iters, int flags,
                        float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* XVIEW_B200_MEASURE_H_ */
