// bin.cu — K3a: tile binning for the compositor.
//
// Part of the replacement of renderer.render_with_pass (reference src/tab/scene.rs:2302-2314):
// the reference draws one instanced quad per visible Gaussian in sorted order and lets the
// rasteriser find the covered pixels; here every depth-sorted splat is expanded into one
// (tile id, splat id) entry per 16x16 tile its extent square touches.  Entries are produced
// IN DEPTH ORDER (order-preserving expansion: block scan + decoupled look-back), so a STABLE
// sort by tile id alone (2 onesweep passes over 16 bits, sort.cu) yields per-tile lists that are
// still front-to-back.  Models are expanded nearest first, each appended after the previous
// one, which reproduces the reference's per-model layering (scene.rs:533-558).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kIpt = 4;                      // depth ranks per thread
constexpr int kChunk = kThreads * kIpt;      // 1024 ranks per chunk
constexpr uint32_t kBigSplat = 32;           // > this many tiles: expanded by the whole warp

struct TileRect { uint32_t tx0, ty0, nx, ny; };

// Pixel bounds of a splat: the integer pixels of the screen-aligned square of half-size
// `radius` around (mx,my), clipped to the viewport — same expression as the compositor's.
__device__ __forceinline__ TileRect tile_rect(float mx, float my, uint32_t radius, float W, float H) {
    TileRect t = {0, 0, 0, 0};
    if (radius == 0) return t;
    float r = (float)radius;
    float fx0 = ceilf(mx - r), fx1 = floorf(mx + r), fy0 = ceilf(my - r), fy1 = floorf(my + r);
    if (fx0 < 0.0f) fx0 = 0.0f;
    if (fy0 < 0.0f) fy0 = 0.0f;
    if (fx1 > W - 1.0f) fx1 = W - 1.0f;
    if (fy1 > H - 1.0f) fy1 = H - 1.0f;
    if (!(fx0 <= fx1 && fy0 <= fy1)) return t;
    t.tx0 = (uint32_t)fx0 / GS_TILE;
    t.ty0 = (uint32_t)fy0 / GS_TILE;
    t.nx = (uint32_t)fx1 / GS_TILE - t.tx0 + 1;
    t.ny = (uint32_t)fy1 / GS_TILE - t.ty0 + 1;
    return t;
}

__global__ void __launch_bounds__(kThreads) k_bin_expand(const uint32_t* __restrict__ sorted_slot,
                                                         const b200gs_splat* __restrict__ splats,
                                                         const uint32_t* d_v, uint32_t v_max, uint32_t splat_base,
                                                         uint64_t* lookback, uint32_t epoch, uint32_t* ticket,
                                                         const uint32_t* entry_base_in, uint32_t* entry_total_out,
                                                         uint32_t* overflow, uint32_t* __restrict__ tile_keys,
                                                         uint32_t* __restrict__ tile_vals, uint32_t capacity,
                                                         float W, float H, uint32_t tiles_x) {
    __shared__ uint32_t s_wsum[kThreads / 32];
    __shared__ uint32_t s_chunk, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v = *d_v;
    if (v > v_max) v = v_max;
    const uint32_t nchunks = (v + kChunk - 1) / kChunk;
    const uint32_t ebase = *entry_base_in;
    if (v == 0) {
        if (blockIdx.x == 0 && tid == 0) *entry_total_out = ebase;
        return;
    }

    while (true) {
        if (tid == 0) s_chunk = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t c = s_chunk;
        if (c >= nchunks) break;

        const uint32_t r0 = c * kChunk + tid * kIpt;
        TileRect tr[kIpt];
        uint32_t id[kIpt], cnt[kIpt], sum = 0;
#pragma unroll
        for (int k = 0; k < kIpt; k++) {
            uint32_t r = r0 + k;
            cnt[k] = 0;
            id[k] = 0;
            tr[k] = TileRect{0, 0, 0, 0};
            if (r < v) {
                uint32_t slot = sorted_slot ? sorted_slot[r] : r;
                uint4 q0 = *reinterpret_cast<const uint4*>(splats + slot);
                tr[k] = tile_rect(__uint_as_float(q0.x), __uint_as_float(q0.y), q0.z & 0xffffu, W, H);
                cnt[k] = tr[k].nx * tr[k].ny;
                id[k] = splat_base + slot;
            }
            sum += cnt[k];
        }
        // block exclusive scan of `sum`
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = lane < kThreads / 32 ? s_wsum[lane] : 0u;
            uint32_t tot = ws;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            tot = __shfl_sync(0xffffffffu, tot, 0);
            uint32_t excl = gs_lookback_warp(lookback, epoch, c, tot, lane);
            if (lane == 0) {
                s_base = excl;
                if (c == nchunks - 1) {
                    uint32_t total = ebase + excl + tot;
                    if (total > capacity) { *overflow = 1u; total = capacity; }
                    *entry_total_out = total;
                }
            }
        }
        __syncthreads();
        uint32_t off = ebase + s_base + incl - sum;
        for (int k = 0; k < warp; k++) off += s_wsum[k];

        // small splats: the owning thread writes its run; big ones are spread over the warp
#pragma unroll
        for (int k = 0; k < kIpt; k++) {
            const bool big = cnt[k] > kBigSplat;
            if (!big) {
                uint32_t o = off;
                for (uint32_t y = 0; y < tr[k].ny; y++)
                    for (uint32_t x = 0; x < tr[k].nx; x++, o++)
                        if (o < capacity) {
                            tile_keys[o] = (tr[k].ty0 + y) * tiles_x + tr[k].tx0 + x;
                            tile_vals[o] = id[k];
                        }
            }
            uint32_t bigmask = __ballot_sync(0xffffffffu, big);
            while (bigmask) {
                int src = __ffs((int)bigmask) - 1;
                bigmask &= bigmask - 1;
                uint32_t b_off = __shfl_sync(0xffffffffu, off, src);
                uint32_t b_cnt = __shfl_sync(0xffffffffu, cnt[k], src);
                uint32_t b_nx = __shfl_sync(0xffffffffu, tr[k].nx, src);
                uint32_t b_tx0 = __shfl_sync(0xffffffffu, tr[k].tx0, src);
                uint32_t b_ty0 = __shfl_sync(0xffffffffu, tr[k].ty0, src);
                uint32_t b_id = __shfl_sync(0xffffffffu, id[k], src);
                for (uint32_t e = lane; e < b_cnt; e += 32) {
                    uint32_t o = b_off + e;
                    if (o < capacity) {
                        uint32_t y = e / b_nx, x = e - y * b_nx;
                        tile_keys[o] = (b_ty0 + y) * tiles_x + b_tx0 + x;
                        tile_vals[o] = b_id;
                    }
                }
            }
            off += cnt[k];
        }
        __syncthreads();  // s_chunk / s_wsum / s_base reused
    }
}

// ranges[tile] = first entry, ranges[n_tiles + tile] = one past the last entry (both 0 if none)
__global__ void __launch_bounds__(256) k_tile_ranges(const uint32_t* __restrict__ tile_keys, const uint32_t* d_entries,
                                                     uint32_t capacity, uint32_t* ranges, uint32_t n_tiles) {
    uint32_t n = *d_entries;
    if (n > capacity) n = capacity;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        uint32_t k = tile_keys[e];
        if (k >= n_tiles) continue;
        if (e == 0 || tile_keys[e - 1] != k) ranges[k] = e;
        if (e == n - 1 || tile_keys[e + 1] != k) ranges[n_tiles + k] = e + 1;
    }
}

}  // namespace

cudaError_t gs_launch_bin(const GsBinArgs& a, const GsFrame& f, int num_sms, cudaStream_t st) {
    static int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_bin_expand, kThreads, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    uint32_t nchunks = (a.v_max + kChunk - 1) / kChunk;
    uint32_t grid = (uint32_t)(blocks_per_sm * num_sms);
    if (grid > nchunks) grid = nchunks;
    if (grid < 1) grid = 1;
    k_bin_expand<<<grid, kThreads, 0, st>>>(a.sorted_slot, a.splats, a.d_v, a.v_max, a.splat_base, a.lookback, a.epoch,
                                            a.ticket, a.entry_base_in, a.entry_total_out, a.overflow, a.tile_keys,
                                            a.tile_vals, a.capacity, f.W, f.H, f.tiles_x);
    return cudaGetLastError();
}

cudaError_t gs_launch_tile_ranges(const uint32_t* tile_keys, const uint32_t* d_entries, uint32_t capacity,
                                  uint32_t* ranges, uint32_t n_tiles, int num_sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)n_tiles * 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_tile_ranges<<<num_sms * 8, 256, 0, st>>>(tile_keys, d_entries, capacity, ranges, n_tiles);
    return cudaGetLastError();
}
