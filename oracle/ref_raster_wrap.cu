// oracle/ref_raster_wrap.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C wrapper around the UNMODIFIED reference rasterizer
// (CudaRasterizer::Rasterizer::forward, cuda_rasterizer/rasterizer.h:31-54 of
// third-party/diff-gaussian-rasterization-w-depth), compiled by oracle/Makefile
// straight from /root/reference into oracle/_ref/libref_raster.so.  It plays the
// role of the reference's torch glue (rasterize_points.cu:35-117): three
// growable scratch buffers handed to the rasterizer through std::function
// callbacks, outputs written into caller-owned device buffers.  Used (a) to pin
// oracle/raster_ref.c and the CUDA kernels against the real reference on the
// GPU box and (b) as the reference-CUDA-rasterizer timing baseline.
#include <cuda_runtime.h>
#include <cstdio>
#include <functional>
#include "rasterizer.h"
#include "rasterizer_impl.h"

namespace {
struct Grow {
    char* ptr = nullptr;
    size_t cap = 0;
    char* get(size_t n)
    {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            size_t want = n + n / 2 + 256;
            if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; return nullptr; }
            cap = want;
        }
        return ptr;
    }
};
Grow g_geom, g_bin, g_img;
}  // namespace

extern "C" {

// All pointers are device pointers (NULL where the reference accepts nullptr).
// Returns num_rendered, or a negative CUDA error code.
int ref_raster_forward(int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                       const float* shs, const float* colors_precomp, const float* opacities,
                       const float* scales, float scale_modifier, const float* rotations,
                       const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                       const float* campos, float tanfovx, float tanfovy, int prefiltered,
                       float z_threshold, float* out_color, float* out_depth, int* radii)
{
    std::function<char*(size_t)> geomF = [](size_t n) { return g_geom.get(n); };
    std::function<char*(size_t)> binF = [](size_t n) { return g_bin.get(n); };
    std::function<char*(size_t)> imgF = [](size_t n) { return g_img.get(n); };
    int rendered = 0;
    if (P != 0) {
        rendered = CudaRasterizer::Rasterizer::forward(
            geomF, binF, imgF, P, D, M, bg, W, H, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx,
            tanfovy, prefiltered != 0, z_threshold, out_color, out_depth, radii);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -(int)e;
    return rendered;
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e;
}

// The reference's own sorted instance list and per-tile ranges of the LAST forward call, located in its scratch
// buffers with its own BinningState / ImageState::fromChunk (rasterizer_impl.cu:172-196), copied device to
// device: point_list_out [num_rendered] Gaussian ids in (tile, depth) order, ranges_out [tiles][2].
int ref_raster_lists(int num_rendered, int W, int H, unsigned* point_list_out, unsigned* ranges_out)
{
    if (!g_bin.ptr || !g_img.ptr) return -1;
    char* bchunk = g_bin.ptr;
    CudaRasterizer::BinningState bin = CudaRasterizer::BinningState::fromChunk(bchunk, (size_t)num_rendered);
    char* ichunk = g_img.ptr;
    CudaRasterizer::ImageState img = CudaRasterizer::ImageState::fromChunk(ichunk, (size_t)W * H);
    const int tiles = ((W + 15) / 16) * ((H + 15) / 16);
    cudaError_t e = cudaSuccess;
    if (num_rendered > 0)
        e = cudaMemcpy(point_list_out, bin.point_list, sizeof(unsigned) * (size_t)num_rendered, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(ranges_out, img.ranges, sizeof(uint2) * (size_t)tiles, cudaMemcpyDeviceToDevice);
    return e == cudaSuccess ? 0 : -(int)e;
}

void ref_raster_release()
{
    if (g_geom.ptr) cudaFree(g_geom.ptr);
    if (g_bin.ptr) cudaFree(g_bin.ptr);
    if (g_img.ptr) cudaFree(g_img.ptr);
    g_geom = Grow(); g_bin = Grow(); g_img = Grow();
}

}  // extern "C"
