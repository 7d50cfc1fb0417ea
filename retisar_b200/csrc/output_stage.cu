// Output stage of a JACK client for S batched streams (reference: JackClient._process_deliver,
// ReTiSAR/_jack_client.py:770-867, with tools.calculate_rms / calculate_peak, tools.py:1117-1169):
// in-place output volume (global x per-channel relative), RMS and peak level per output channel,
// clipping count.  One warp per (stream, channel) row; the row is read once, scaled, written back
// and reduced in registers -- an HBM-bound pass of 8 bytes per sample.
#include "rtb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
output_stage_kernel(float *__restrict__ y, const float *__restrict__ gain, float *__restrict__ rms,
                    float *__restrict__ peak, unsigned int *__restrict__ clip_count, long long rows,
                    int n_ch, int B, int as_level) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float g = gain[row % n_ch];
    float4 *p = reinterpret_cast<float4 *>(y + row * B);
    float ss = 0.f, mx = 0.f;
    for (int i = lane; i < B / 4; i += 32) {
        float4 v = p[i];
        v.x *= g; v.y *= g; v.z *= g; v.w *= g;
        p[i] = v;
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        mx = fmaxf(fmaxf(fmaxf(mx, fabsf(v.x)), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) {
        float r = sqrtf(ss / (float)B), pk = mx;
        if (clip_count && mx > 1.0f) atomicAdd(clip_count, 1u);  // `_is_detect_clipping`: peak > 1 (:856)
        if (as_level) {  // 20 log10, zeros reported as -200 dB (tools.py:1141-1145, :1165-1169)
            r = r == 0.f ? -200.f : 20.f * log10f(r);
            pk = pk == 0.f ? -200.f : 20.f * log10f(pk);
        }
        if (rms) rms[row] = r;
        if (peak) peak[row] = pk;
    }
}

}  // namespace

extern "C" int rtb_output_stage(float *d_out, const float *d_gain, float *d_rms, float *d_peak,
                                unsigned int *d_clip_count, long long n_streams, int n_channels,
                                int block_length, int as_level, void *cuda_stream) {
    if (!d_out || !d_gain) return rtb::fail(RTB_ERR_INVALID, "NULL argument");  // d_clip_count NULL: detection off
    if (n_streams < 1 || n_channels < 1 || block_length < 4 || block_length % 4)
        return rtb::fail(RTB_ERR_INVALID, "bad sizes (block length must be a multiple of 4)");
    const long long rows = n_streams * n_channels;
    output_stage_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)cuda_stream>>>(
        d_out, d_gain, d_rms, d_peak, d_clip_count, rows, n_channels, block_length, as_level);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
