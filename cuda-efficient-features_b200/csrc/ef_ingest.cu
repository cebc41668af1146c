// ef_ingest.cu -- the step right before the path: colour frames -> CV_8UC1.
// Replaces convertToGray (samples/sample_common.cpp:35-45: cv::cvtColor BGR2GRAY / BGRA2GRAY on the host) with a device kernel
// so that a colour frame is uploaded once and never round-trips.  Arithmetic = OpenCV's 8-bit path (third-party, imgproc
// color_rgb: fixed point, 15 fractional bits): gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15; pinned against cv2 4.13 in
// tests/test_matcher_cpu.py (all 2^24 colours).
#include "ef_common.cuh"

// one thread = 4 output pixels: 12 (BGR) or 16 (BGRA) input bytes, one 32-bit store
template <int CN>
__global__ void __launch_bounds__(256) ef_bgr_to_gray_kernel(const uint8_t* __restrict__ src, size_t spitch, int w, int h, uint8_t* __restrict__ dst, size_t dpitch)
{
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= w) return;
    const uint8_t* sp = src + (size_t)y * spitch + (size_t)x0 * CN;
    uint8_t* dp = dst + (size_t)y * dpitch + x0;
    unsigned px[4 * CN / 4 + 1];
    const bool fast = x0 + 3 < w && ((reinterpret_cast<uintptr_t>(sp) & 3) == 0);
    unsigned g[4];
    if (fast) {
#pragma unroll
        for (int i = 0; i < CN; i++) px[i] = reinterpret_cast<const unsigned*>(sp)[i];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int o = i * CN;
            const unsigned b = (px[o >> 2] >> (8 * (o & 3))) & 0xffu;
            const unsigned gr = (px[(o + 1) >> 2] >> (8 * ((o + 1) & 3))) & 0xffu;
            const unsigned r = (px[(o + 2) >> 2] >> (8 * ((o + 2) & 3))) & 0xffu;
            g[i] = (b * 3735u + gr * 19235u + r * 9798u + (1u << 14)) >> 15;
        }
        if ((reinterpret_cast<uintptr_t>(dp) & 3) == 0) { *reinterpret_cast<unsigned*>(dp) = g[0] | (g[1] << 8) | (g[2] << 16) | (g[3] << 24); return; }
        for (int i = 0; i < 4; i++) dp[i] = (uint8_t)g[i];
        return;
    }
    for (int i = 0; i < 4 && x0 + i < w; i++) {
        const unsigned b = sp[i * CN], gr = sp[i * CN + 1], r = sp[i * CN + 2];
        dp[i] = (uint8_t)((b * 3735u + gr * 19235u + r * 9798u + (1u << 14)) >> 15);
    }
}

extern "C" int ef_bgr_to_gray_async(const uint8_t* d_src, size_t src_pitch, int width, int height, int channels,
                                    uint8_t* d_gray, size_t gray_pitch, void* stream)
{
    if (!d_src || !d_gray || width <= 0 || height <= 0) return EF_ERR_BAD_ARG;
    if (channels != 3 && channels != 4) return EF_ERR_BAD_ARG;   // CV_Error(StsBadArg, "Image should be 8UC1, 8UC3 or 8UC4")
    if (src_pitch < (size_t)width * channels || gray_pitch < (size_t)width) return EF_ERR_BAD_ARG;
    const dim3 grid(ef_div_up(ef_div_up(width, 4), 256), height);
    if (channels == 3) ef_bgr_to_gray_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, src_pitch, width, height, d_gray, gray_pitch);
    else ef_bgr_to_gray_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, src_pitch, width, height, d_gray, gray_pitch);
    EF_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? EF_OK : EF_ERR_CUDA;
}

// ---- synthetic frames (measurement support; SURVEY 8d "pinned counter-based generator"): pix(f, y, x) = lowbias32(seed ^ ((f * H + y) * W + x)) >> 24,
// the generator of the CPU arm (oracle efo_synth_frame) restated on the device so that bench.py feeds BOTH arms the same pixels
// without an upload.  One thread = 4 adjacent pixels, one 32-bit store when aligned.
__device__ __forceinline__ unsigned ef_lowbias32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(256) ef_synth_frames_kernel(uint8_t* __restrict__ dst, size_t pitch, size_t stride, int w, int h, unsigned seed, unsigned first_frame)
{
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= w) return;
    const unsigned frame = first_frame + blockIdx.z;
    const unsigned base = (frame * (unsigned)h + (unsigned)y) * (unsigned)w + (unsigned)x0;
    uint8_t* dp = dst + (size_t)blockIdx.z * stride + (size_t)y * pitch + x0;
    unsigned v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = ef_lowbias32(seed ^ (base + i)) >> 24;
    if (x0 + 3 < w && (reinterpret_cast<uintptr_t>(dp) & 3) == 0) *reinterpret_cast<unsigned*>(dp) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    else for (int i = 0; i < 4 && x0 + i < w; i++) dp[i] = (uint8_t)v[i];
}

extern "C" int ef_synth_frames_async(uint8_t* d_frames, size_t pitch, size_t frame_stride, int width, int height, int nframes,
                                     unsigned seed, unsigned first_frame, void* stream)
{
    if (!d_frames || width <= 0 || height <= 0 || nframes <= 0 || nframes > 65535 || pitch < (size_t)width) return EF_ERR_BAD_ARG;
    const dim3 grid(ef_div_up(ef_div_up(width, 4), 256), height, nframes);
    ef_synth_frames_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_frames, pitch, frame_stride, width, height, seed, first_frame);
    EF_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? EF_OK : EF_ERR_CUDA;
}
