// stand-alone TMA probe: tma_probe <variant>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned bytes, unsigned* out)
{
    extern __shared__ __align__(1024) unsigned char s_in[];
    __shared__ __align__(8) unsigned long long s_mbar;
    const unsigned mbar = smem_addr(&s_mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_addr(s_in)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_addr(s_in)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(mbar) : "memory");
    }
    unsigned done = 0;
    for (unsigned spin = 0; !done && spin < (1u << 22); spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
    if (threadIdx.x == 0) out[0] = done;
    for (unsigned i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[1 + i] = reinterpret_cast<unsigned*>(s_in)[i];
}

__constant__ CUtensorMap c_map;
template <int WHERE>   // 0: global memory pointer, 1: __constant__
__global__ void k2(const CUtensorMap* gmap, int x, int y, unsigned bytes, unsigned* out)
{
    extern __shared__ __align__(1024) unsigned char s_in[];
    __shared__ __align__(8) unsigned long long s_mbar;
    const unsigned mbar = smem_addr(&s_mbar);
    const CUtensorMap* mp = WHERE == 0 ? gmap : &c_map;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_addr(s_in)), "l"(reinterpret_cast<unsigned long long>(mp)), "r"(x), "r"(y), "r"(mbar) : "memory");
    }
    unsigned done = 0;
    for (unsigned spin = 0; !done && spin < (1u << 22); spin++)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
    if (threadIdx.x == 0) out[0] = done;
    for (unsigned i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[1 + i] = reinterpret_cast<unsigned*>(s_in)[i];
}

int main(int argc, char** argv)
{
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    printf("variant %d driver %d runtime %d\n", v, drv, rt);
    const int w = 400, h = 300, nf = 2; const size_t pitch = 400, stride = pitch * h;
    std::vector<unsigned char> img(stride * nf);
    for (size_t i = 0; i < img.size(); i++) img[i] = (unsigned char)(i * 7 + i / 400);
    unsigned char* d_img; cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fp=%p\n", cudaGetErrorString(ge), (int)q, fp);
    EncodeTiledFn fn = (EncodeTiledFn)fp;
    int rank = (v == 1 || v == 3 || v >= 5) ? 2 : 3;
    cuuint32_t bw = (v == 2) ? 128 : ((v == 3 || v >= 5) ? 64 : 80), bh = (v == 2 || v == 3 || v >= 5) ? 64 : 70;
    alignas(64) CUtensorMap map; memset(&map, 0, sizeof(map));
    const cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nf };
    const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)stride };
    const cuuint32_t box[3] = { bw, bh, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    v == 4 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rank %d box %ux%u: %d\n", rank, bw, bh, (int)r);
    if (argc > 2 && atoi(argv[2]) == 1) reinterpret_cast<unsigned long long*>(&map)[1] &= ~(1ull << 21);
    const unsigned long long* mw = (const unsigned long long*)&map;
    for (int i = 0; i < 16; i++) printf("%016llx%c", mw[i], i % 4 == 3 ? '\n' : ' ');
    const unsigned bytes = bw * bh;
    unsigned* d_out; cudaMalloc(&d_out, bytes + 4); cudaMemset(d_out, 0, bytes + 4);
    const int x = argc > 3 ? atoi(argv[3]) : 56, y = 61, z = 1;
    if (v == 5) { CUtensorMap* dm; cudaMalloc(&dm, sizeof(map)); cudaMemcpy(dm, &map, sizeof(map), cudaMemcpyHostToDevice); k2<0><<<1, 128, bytes>>>(dm, x, y, bytes, d_out); }
    else if (v == 6) { cudaMemcpyToSymbol(c_map, &map, sizeof(map)); k2<1><<<1, 128, bytes>>>(nullptr, x, y, bytes, d_out); }
    else if (rank == 3) k<3><<<1, 128, bytes>>>(map, x, y, z, bytes, d_out); else k<2><<<1, 128, bytes>>>(map, x, y, z, bytes, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<unsigned> out(bytes / 4 + 1);
    cudaMemcpy(out.data(), d_out, bytes + 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (unsigned r2 = 0; r2 < bh; r2++) for (unsigned c = 0; c < bw; c++) {
        const int gx = x + c, gy = y + r2, zz = rank == 3 ? z : 0;
        unsigned char want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? img[zz * stride + gy * pitch + gx] : 0;
        bad += want != ((unsigned char*)(out.data() + 1))[r2 * bw + c];
    }
    printf("done flag %u mismatches %d\n", out[0], bad);
    return 0;
}
