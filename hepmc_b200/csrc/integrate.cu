// Plain MC and VEGAS importance-sampling integration, fused: Philox draw -> grid map ->
// integrand -> weight -> per-thread shifted moments + per-(axis,bin) histogram.
//
// Determinism / GPU-count invariance: the global point range of one iteration is cut into
// fixed CHUNKS (one CTA each, fixed point->thread->order mapping); every chunk writes one
// partial row (count, mean, M2 | hist[ndim*K]); partial rows are merged in chunk order by
// the reduce kernels.  No floating-point atomics anywhere, so results are bit-identical for
// any partition of the chunk range over GPUs.
//
// Histogram without atomics: after each warp-tile of 32 points the lanes swap roles -- lane
// L owns axis d = L % ndim and a PRIVATE column hist[k][L] in shared memory, and walks the
// tile's points in fixed order (plain LDS/DADD/STS, conflict-free banks).  The per-lane
// columns are summed in fixed order when the chunk ends.
#include "integrate_kernels.cuh"

namespace hepmc {

// ------------------------------------------------------------------ two-kernel path: accumulate
__global__ void __launch_bounds__(256) vegas_accumulate_kernel(const __grid_constant__ SampleArgs a) {
    extern __shared__ double smem_raw[];
    const int ndim = a.ndim, K = a.K;
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SmemLayout s = carve(smem_raw, 0, K, nw);  // no grid needed
    for (int i = threadIdx.x; i < nw * K * 32; i += blockDim.x) s.hist[i] = 0.0;
    __syncthreads();
    double* hist_w = s.hist + (size_t)warp * K * 32;
    double* wf_w = s.wf + warp * 32;
    uint32_t* bins_w = s.bins + warp * 32 * kBinWords;
    const int64_t chunk0 = (int64_t)blockIdx.x * a.chunk;
    const int ppt = (int)(a.chunk / blockDim.x);
    ShiftedAcc acc;
    acc.init();
    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        const bool valid = li < a.n;
        double wf = 0.0;
        for (int q = 0; q * 4 < ndim; ++q) {
            uint32_t word = 0u;
            if (valid)
                for (int b = 0; b < 4 && q * 4 + b < ndim; ++b)
                    word |= ((uint32_t)a.idx_in[(int64_t)(q * 4 + b) * a.n + li] & 0xffu) << (8 * b);
            bins_w[lane * kBinWords + q] = word;
        }
        if (valid) {
            wf = a.f_in[li] * a.w_in[li];
            acc.add(wf);
        }
        wf_w[lane] = wf;
        __syncwarp();
        warp_hist_update(hist_w, wf_w, bins_w, ndim, lane);
        __syncwarp();
    }
    __syncthreads();
    block_hist_flush(s.hist, ndim, K, nw, a.hist_partials + (size_t)blockIdx.x * ndim * K);
    // full chunk: equal counts in every lane -> the division-free butterfly (same bits, common.cuh)
    Stats st = (chunk0 + a.chunk <= a.n) ? warp_merge_equal(acc.stats()) : warp_merge(acc.stats());
    if (lane == 0) s.stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < nw; ++w) r = merge(r, s.stats[w]);
        double* o = a.stats_partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// values (optionally / weights) -> chunk partials; importance.py:47-50
__global__ void __launch_bounds__(256) stats_partials_kernel(const double* __restrict__ values,
                                                             const double* __restrict__ weights, double wscalar,
                                                             int64_t n, int64_t chunk, double* __restrict__ partials) {
    __shared__ Stats s_stats[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunk0 = (int64_t)blockIdx.x * chunk;
    const int ppt = (int)(chunk / blockDim.x);
    ShiftedAcc acc;
    acc.init();
    for (int it = 0; it < ppt; ++it) {
        const int64_t li = chunk0 + (int64_t)it * blockDim.x + threadIdx.x;
        if (li >= n) break;
        double v = values[li];
        if (weights) v = v / weights[li];
        else if (wscalar != 1.0) v = v / wscalar;
        acc.add(v);
    }
    // full chunk: equal counts in every lane -> the division-free butterfly (same bits, common.cuh)
    Stats st = (chunk0 + chunk <= n) ? warp_merge_equal(acc.stats()) : warp_merge(acc.stats());
    if (lane == 0) s_stats[warp] = st;
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats r{0.0, 0.0, 0.0};
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = merge(r, s_stats[w]);
        double* o = partials + (size_t)blockIdx.x * 3;
        o[0] = r.n; o[1] = r.mean; o[2] = r.m2;
    }
}

// ------------------------------------------------------------------ ordered reductions
constexpr int kMaxOut = 64;
struct RowRanges {
    int64_t begin[kMaxOut + 1];
    int n_out;
};

// one CTA per output slot: thread t merges the rows b + t, b + t + 256, ... in row order, then the 256
// thread results are merged in a fixed tree (warp butterfly, lower lane on the left; the 8 warp
// results in warp order).  The tree depends on the slot's row range only, never on the GPU count.
// (Round 1 used ONE warp per slot: 119 dependent merges per lane plus a 32-step serial lane merge,
// 80 us per iteration -- the largest fixed cost of the 8-GPU step.)
__global__ void __launch_bounds__(256) reduce_stats_kernel(const double* __restrict__ rows,
                                                           const __grid_constant__ RowRanges rr,
                                                           double* __restrict__ out, int out_stride) {
    __shared__ Stats s_w[8];
    const int s = blockIdx.x;
    const int64_t b = rr.begin[s], e = rr.begin[s + 1];
    Stats acc{0.0, 0.0, 0.0};
    for (int64_t r = b + threadIdx.x; r < e; r += 256) {
        Stats t{rows[r * 3], rows[r * 3 + 1], rows[r * 3 + 2]};
        acc = merge(acc, t);
    }
    const Stats total = block_merge<256>(acc, s_w);
    if (threadIdx.x == 0) {
        double* o = out + (size_t)s * out_stride;
        o[0] = total.n;
        o[1] = total.mean;
        o[2] = total.m2;
    }
}

// out[s][col] = sum over the rows of slot s.  kHistLanes row lanes per column: lane y adds rows
// b + y, b + y + kHistLanes, ... in row order, the lane sums are combined pairwise in a fixed
// tree.  The order depends on the slot's row range only (never on the GPU count), and the
// dependent add chain per thread is rows / 16 long: the one-thread-per-column version took
// 0.39 ms per iteration at 1e9 points -- 1 % of the step on one GPU but 7 % on eight, where a
// rank's slot is as long as before while everything else shrinks.
constexpr int kHistLanes = 16;
constexpr int kHistCols = 32;

__global__ void __launch_bounds__(kHistLanes * kHistCols) reduce_hist_kernel(const double* __restrict__ rows,
                                                                             const __grid_constant__ RowRanges rr,
                                                                             int ncols, double* __restrict__ out) {
    __shared__ double part[kHistLanes][kHistCols];
    const int s = blockIdx.y;
    const int cx = threadIdx.x % kHistCols, ry = threadIdx.x / kHistCols;
    const int col = blockIdx.x * kHistCols + cx;
    const int64_t b = rr.begin[s], e = rr.begin[s + 1];
    double acc = 0.0;
    if (col < ncols) {
        for (int64_t r = b + ry; r < e; r += kHistLanes) acc += rows[r * ncols + col];
    }
    part[ry][cx] = acc;
    __syncthreads();
#pragma unroll
    for (int half = kHistLanes / 2; half > 0; half >>= 1) {
        if (ry < half) part[ry][cx] += part[ry + half][cx];
        __syncthreads();
    }
    if (ry == 0 && col < ncols) out[(size_t)s * ncols + col] = part[0][cx];
}

// ------------------------------------------------------------------ fused slot reduction (VEGAS iteration)
// Chunk partial rows (count, mean, M2 | hist[cols]) of every slot -> one PACKED row per slot
// [count, mean, M2, hist...] (the all-gather send buffer), in two levels so that one slot's rows (3815
// at 1e9 points) are read by kSlotGroups * (column tiles + 1) CTAs instead of 8:
//   level 1: group g of slot s owns the rows [b + len*g/G, b + len*(g+1)/G); histogram CTAs add them per
//            column with 8 row lanes (rows r0+y, r0+y+8, ... in order, lanes combined in a fixed tree);
//            the statistics CTA Chan-merges the group's rows (thread order, fixed tree);
//   level 2: the G group partials of a slot are combined in group order.
// Every order is a function of the slot's row range only: bit-identical at any GPU count.
constexpr int kSlotGroups = 32;
constexpr int kRedCols = 32, kRedLanes = 8;

// level 2 of slot s, run by the LAST level-1 CTA of the slot to finish (threadfence + counter): one launch,
// and the combine order is fixed whichever CTA runs it.
__device__ __forceinline__ void reduce_slots_l2(const double* __restrict__ in /* (G, 3 + ncols) */, int ncols,
                                                double* __restrict__ out /* (3 + ncols) */) {
    static_assert(kSlotGroups == 32, "the group statistics are merged by one warp, one group per lane");
    const int row = 3 + ncols;
    // the group statistics first: lane g of warp 0 loads group g (all loads in flight at once -- a single thread
    // walking the 32 groups paid an L2 round trip per group: 13 us), merged in the fixed butterfly of warp_merge
    if (threadIdx.x < 32) {
        const double* gs = in + (size_t)threadIdx.x * row;
        Stats u{__ldcg(gs), __ldcg(gs + 1), __ldcg(gs + 2)};
        const Stats t = warp_merge(u);
        if (threadIdx.x == 0) { out[0] = t.n; out[1] = t.mean; out[2] = t.m2; }
    }
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
        double v[kSlotGroups];
#pragma unroll
        for (int g = 0; g < kSlotGroups; ++g) v[g] = __ldcg(in + (size_t)g * row + 3 + c);   // 32 independent loads
        double acc = 0.0;
#pragma unroll
        for (int g = 0; g < kSlotGroups; ++g) acc += v[g];                                  // group order
        out[3 + c] = acc;
    }
}

// Peer-memory exchange of the packed slot rows (the all-gather of the VEGAS iteration, fused into the reduction):
// every rank owns a symmetric buffer  [2 parities][HEPMC_N_SLOTS][3 + ncols] doubles | [2][HEPMC_N_SLOTS] flags
// mapped into all peers (NVLink P2P through NVSwitch).  The CTA that finishes a slot writes its packed row
// straight into EVERY rank's buffer at the slot's global position and then publishes flag = epoch there
// (release at system scope); the grid update of every rank waits for its 8 flags (vegas_update_kernel).
// Two parities: a rank can be at most one iteration ahead of a peer (its next update waits for the peer's next
// row), so the buffer a slow peer still reads is never the one being written.
struct SlotPeers {
    double* base[HEPMC_N_SLOTS];     // peer buffers (device pointers valid on this GPU); world <= 8
    int world;
    int slot0;                       // global index of this rank's first slot
    unsigned long long epoch;        // iteration counter, > 0, the same on every rank
};

__device__ __forceinline__ unsigned long long* peer_flags(double* base, int ncols) {
    return reinterpret_cast<unsigned long long*>(base + (size_t)2 * HEPMC_N_SLOTS * (3 + ncols));
}

__global__ void __launch_bounds__(kRedCols * kRedLanes) reduce_slots_kernel(
    const double* __restrict__ stats_rows, const double* __restrict__ hist_rows, const __grid_constant__ RowRanges rr,
    int ncols, double* __restrict__ scratch /* (n_out, G, 3 + ncols) */, unsigned int* __restrict__ counters /* n_out, zero */,
    double* __restrict__ packed /* (n_out, 3 + ncols) */, const __grid_constant__ SlotPeers peers) {
    __shared__ double part[kRedLanes][kRedCols];
    __shared__ Stats s_w[8];
    __shared__ bool s_last;
    const int s = blockIdx.z, g = blockIdx.y;
    const int64_t b = rr.begin[s], len = rr.begin[s + 1] - b;
    const int64_t r0 = b + len * g / kSlotGroups, r1 = b + len * (g + 1) / kSlotGroups;
    double* out = scratch + ((size_t)s * kSlotGroups + g) * (3 + ncols);
    const int ntile = (ncols + kRedCols - 1) / kRedCols;
    if ((int)blockIdx.x == ntile) {   // statistics of the group
        Stats acc{0.0, 0.0, 0.0};
        for (int64_t r = r0 + threadIdx.x; r < r1; r += kRedCols * kRedLanes) {
            Stats t{stats_rows[r * 3], stats_rows[r * 3 + 1], stats_rows[r * 3 + 2]};
            acc = merge(acc, t);
        }
        const Stats total = block_merge<kRedCols * kRedLanes>(acc, s_w);
        if (threadIdx.x == 0) { out[0] = total.n; out[1] = total.mean; out[2] = total.m2; }
    } else {
        const int cx = threadIdx.x % kRedCols, ry = threadIdx.x / kRedCols;
        const int col = blockIdx.x * kRedCols + cx;
        double acc = 0.0;
        if (col < ncols) {
            // eight rows in flight per step, added in row order (one dependent L2 round trip per ROW made this loop
            // the longest part of the reduction: ~15 rows per lane at 1e9 points on 8 GPUs)
            int64_t r = r0 + ry;
            for (; r + 7 * kRedLanes < r1; r += 8 * kRedLanes) {
                double v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __ldcg(hist_rows + (r + (int64_t)i * kRedLanes) * ncols + col);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += v[i];
            }
            for (; r < r1; r += kRedLanes) acc += __ldcg(hist_rows + r * ncols + col);
        }
        part[ry][cx] = acc;
        __syncthreads();
#pragma unroll
        for (int half = kRedLanes / 2; half > 0; half >>= 1) {
            if (ry < half) part[ry][cx] += part[ry + half][cx];
            __syncthreads();
        }
        if (ry == 0 && col < ncols) out[3 + col] = part[0][cx];
    }
    // last CTA of the slot: combine the groups
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = (unsigned int)(ntile + 1) * kSlotGroups;
        const unsigned int prev = atomicAdd(counters + s, 1u);
        s_last = (prev == total - 1);
        if (s_last) counters[s] = 0u;     // ready for the next call on the same scratch
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double* mine = packed + (size_t)s * (3 + ncols);
        reduce_slots_l2(scratch + (size_t)s * kSlotGroups * (3 + ncols), ncols, mine);
        if (peers.world > 1) {
            // publish: the row into every rank's buffer (plain stores over NVLink), then the flag
            __syncthreads();
            const int row = 3 + ncols;
            const size_t at = ((size_t)(peers.epoch & 1) * HEPMC_N_SLOTS + peers.slot0 + s) * row;
            for (int r = 0; r < peers.world; ++r)
                for (int c = threadIdx.x; c < row; c += blockDim.x) peers.base[r][at + c] = mine[c];
            __threadfence_system();
            __syncthreads();
            if ((int)threadIdx.x < peers.world) {
                unsigned long long* f = peer_flags(peers.base[threadIdx.x], ncols) + (peers.epoch & 1) * HEPMC_N_SLOTS + peers.slot0 + s;
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(peers.epoch) : "memory");
            }
        }
    }
}

// Body of the grid adaptation, shared by vegas_update_kernel and the fused small-problem kernel: one CTA of 256
// threads, `sm` = ndim*K doubles of dynamic shared memory.
__device__ __forceinline__ void vegas_update_body(const double* __restrict__ slot_stats, int64_t stats_stride,
                                                  const double* __restrict__ slot_hist, int64_t hist_stride,
                                                  int n_slots, double* __restrict__ grid, int ndim, int K,
                                                  double c, int j, double* __restrict__ est_var, double* sm) {
    double* s_new = sm;  // ndim*K new sizes
    __shared__ double s_est, s_n;
    __shared__ Stats s_slot[64];
    // slot statistics: loaded by n_slots threads at once, merged in slot order by thread 0
    if ((int)threadIdx.x < n_slots && threadIdx.x < 64) {
        const double* ss = slot_stats + (size_t)threadIdx.x * stats_stride;
        s_slot[threadIdx.x] = Stats{__ldcg(ss), __ldcg(ss + 1), __ldcg(ss + 2)};
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        Stats t{0.0, 0.0, 0.0};
        for (int s = 0; s < n_slots; ++s) {
            Stats u;
            if (s < 64) u = s_slot[s];
            else {
                const double* ss = slot_stats + (size_t)s * stats_stride;
                u = Stats{__ldcg(ss), __ldcg(ss + 1), __ldcg(ss + 2)};
            }
            t = merge(t, u);
        }
        s_est = t.mean;               // est_j = mean(weighted)           vegas.py:110
        s_n = t.n;
        est_var[2 * j] = t.mean;
        est_var[2 * j + 1] = t.m2 / t.n / t.n;  // var_j = var0(weighted)/N   vegas.py:111
    }
    __syncthreads();
    const double est = s_est, n = s_n;
    const int stride = 2 * K + 1;
    const double inertia = (double)j;
    for (int q = threadIdx.x; q < ndim * K; q += blockDim.x) {
        const int d = q / K, k = q % K;
        double h = 0.0;
        {   // slot order; the loads of eight slots are independent (a dependent L2 round trip per slot cost ~5 us)
            int s = 0;
            for (; s + 8 <= n_slots; s += 8) {
                double v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __ldcg(slot_hist + (size_t)(s + i) * hist_stride + q);
#pragma unroll
                for (int i = 0; i < 8; ++i) h += v[i];
            }
            for (; s < n_slots; ++s) h += __ldcg(slot_hist + (size_t)s * hist_stride + q);
        }
        const double estimate = h / n;                       // vegas.py:108-109
        const double size = grid[d * stride + K + 1 + k];
        const double old_v = est / size;                     // vegas.py:113
        const double new_v = (double)K * estimate;
        // damped_update, helper.py:71-72
        const double inv = new_v * c / (inertia + 1.0 + c) + old_v * (inertia + 1.0) / (inertia + 1.0 + c);
        s_new[q] = 1.0 / inv;                                // vegas.py:115
    }
    __syncthreads();
    // normalisation in three steps so that the K divisions of an axis run in parallel (one thread per axis doing
    // sum, K divisions and the running sum in one loop was ~5 us of a 10 us kernel at K = 50); same values
    __shared__ double s_tot[HEPMC_MAX_DIM];
    if (threadIdx.x < ndim) {
        const int d = threadIdx.x;
        double tot = 0.0;
        for (int k = 0; k < K; ++k) tot += s_new[d * K + k];
        s_tot[d] = tot;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < ndim * K; q += blockDim.x) {
        const int d = q / K, k = q % K;
        const double sz = s_new[q] / s_tot[d];               // vegas.py:117
        s_new[q] = sz;
        grid[d * stride + K + 1 + k] = sz;
    }
    __syncthreads();
    if (threadIdx.x < ndim) {
        const int d = threadIdx.x;
        double* gd = grid + d * stride;
        double cum = 0.0;
        gd[0] = 0.0;                                         // App. B2: bounds[d][0] = 0
        for (int k = 0; k < K; ++k) {
            cum += s_new[d * K + k];                         // np.cumsum, stratified_volume.py:167
            gd[k + 1] = cum;
        }
    }
}

// ------------------------------------------------------------------ grid adaptation (one CTA)
// slot s: statistics at slot_stats + s*stride, histogram at slot_hist + s*stride (stride 3 / ndim*K for
// separate arrays; 3 + ndim*K for the packed rows of the all-gather buffer)
__global__ void __launch_bounds__(256) vegas_update_kernel(const double* __restrict__ slot_stats, int64_t stats_stride,
                                                           const double* __restrict__ slot_hist, int64_t hist_stride,
                                                           int n_slots, double* __restrict__ grid, int ndim, int K,
                                                           double c, int j, double* __restrict__ est_var,
                                                           const unsigned long long* __restrict__ wait_flags,
                                                           unsigned long long epoch, int* __restrict__ timeout_flag) {
    extern __shared__ double sm[];
    if (wait_flags != nullptr) {
        // peer-memory exchange: slot s is complete when its flag carries this iteration's epoch.  Bounded spin
        // (~4 s): a missing peer becomes an error code of the call, never a hung GPU.
        if ((int)threadIdx.x < n_slots) {
            const long long t0 = clock64();
            unsigned long long v = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(wait_flags + threadIdx.x) : "memory");
                if (v >= epoch) break;
                if (clock64() - t0 > 8000000000LL) { if (timeout_flag) *timeout_flag = 1 + (int)threadIdx.x; break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    vegas_update_body(slot_stats, stats_stride, slot_hist, hist_stride, n_slots, grid, ndim, K, c, j, est_var, sm);
}

// ------------------------------------------------------------------ small problems: reduction + adaptation in one CTA
// When no slot has more than kSlotGroups rows, every group of the two-level slot reduction holds at most one row: its
// level-1 result IS that row (merging with empty statistics and adding +0.0 change nothing), and level 2 combines the
// rows of a slot in group order.  Then one CTA can do the slot reduction and the grid adaptation of a VEGAS iteration
// with the same values, bit for bit, as reduce_slots_kernel + vegas_update_kernel (test_vegas_small_step_*): one launch
// of ~6 us instead of two of 14 + 10 us -- at the reference's own sizes (1e5 points per iteration) the iteration is
// three dependent short kernels and nothing else.
__global__ void __launch_bounds__(256) vegas_small_step_kernel(const double* __restrict__ stats_rows,
                                                               const double* __restrict__ hist_rows,
                                                               const __grid_constant__ RowRanges rr, int ncols,
                                                               double* __restrict__ packed /* (n_out, 3 + ncols) */,
                                                               double* __restrict__ grid, int ndim, int K, double c, int j,
                                                               double* __restrict__ est_var) {
    extern __shared__ double sm[];
    const int row = 3 + ncols;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // statistics: lane g holds the row of group g (or empty statistics), merged in warp_merge's fixed butterfly --
    // what reduce_slots_l2 does with the level-1 results
    for (int s = warp; s < rr.n_out; s += 8) {
        const int64_t b = rr.begin[s], len = rr.begin[s + 1] - b;
        const int64_t r0 = len * lane / kSlotGroups, r1 = len * (lane + 1) / kSlotGroups;
        Stats u{0.0, 0.0, 0.0};
        if (r1 > r0) {
            const double* p = stats_rows + (b + r0) * 3;
            u = Stats{p[0], p[1], p[2]};
        }
        const Stats t = warp_merge(u);
        if (lane == 0) {
            double* o = packed + (size_t)s * row;
            o[0] = t.n; o[1] = t.mean; o[2] = t.m2;
        }
    }
    // histograms: rows of the slot in order
    for (int idx = threadIdx.x; idx < rr.n_out * ncols; idx += blockDim.x) {
        const int s = idx / ncols, col = idx - s * ncols;
        const int64_t b = rr.begin[s], e = rr.begin[s + 1];
        double acc = 0.0;
        int64_t r = b;
        for (; r + 8 <= e; r += 8) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = hist_rows[(r + i) * ncols + col];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += v[i];
        }
        for (; r < e; ++r) acc += hist_rows[r * ncols + col];
        packed[(size_t)s * row + 3 + col] = acc;
    }
    __threadfence();
    __syncthreads();
    vegas_update_body(packed, row, packed + 3, row, rr.n_out, grid, ndim, K, c, j, est_var, sm);
}

// ------------------------------------------------------------------ GridVolumes.pdf
__global__ void __launch_bounds__(256) grid_pdf_kernel(const double* __restrict__ grid, int ndim, int K, double kpow,
                                                       const double* __restrict__ xs, int64_t n,
                                                       double* __restrict__ out) {
    extern __shared__ double s_grid[];
    const int stride = 2 * K + 1;
    for (int i = threadIdx.x; i < ndim * stride; i += blockDim.x) s_grid[i] = grid[i];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool inside = true;
        double vol = 1.0;
        for (int d = 0; d < ndim; ++d) {
            const double x = xs[i * ndim + d];
            inside = inside && (0.0 < x) && (x < 1.0);
            const double* b = s_grid + d * stride;
            // np.argmax(x < bounds) - 1 (stratified_volume.py:80-81); all-False -> -1 -> last bin
            int idx = -1;
            for (int k = 0; k <= K; ++k)
                if (x < b[k]) { idx = k - 1; break; }
            if (idx < 0) idx = K - 1;
            vol *= b[K + 1 + idx];
        }
        out[i] = inside ? 1.0 / kpow / vol : 0.0;  // pdf_indices :67-72 with default counts
    }
}

static int fill_ranges(RowRanges& rr, const int64_t* row_begin_host, int n_out, const char* who) {
    HEPMC_REQUIRE(row_begin_host != nullptr && n_out >= 1 && n_out <= kMaxOut, HEPMC_E_ARG,
                  "%s: n_out=%d outside [1,%d] or NULL ranges", who, n_out, kMaxOut);
    rr.n_out = n_out;
    for (int i = 0; i <= n_out; ++i) rr.begin[i] = row_begin_host[i];
    for (int i = 0; i < n_out; ++i)
        HEPMC_REQUIRE(rr.begin[i] <= rr.begin[i + 1], HEPMC_E_ARG, "%s: row ranges must be non-decreasing", who);
    return 0;
}

static int pick_threads(int ndim, int K, size_t& smem) {
    int dev = 0, maxs = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    for (int t = 256; t >= 64; t >>= 1) {
        smem = vegas_smem_bytes(ndim, K, t);
        if ((int64_t)smem <= (int64_t)maxs) return t;
    }
    return 0;
}

// aligned = false: pure draw generation (no integrand, no partial rows), where the chunk grid carries no
// reduction order and first_index may be any global index (per-chain draws of the generic chain path)
static int check_sample_common(const char* who, int ndim, int K, int64_t n, int64_t first_index, int64_t chunk,
                               bool aligned = true) {
    HEPMC_REQUIRE(ndim >= 1 && ndim <= HEPMC_MAX_DIM, HEPMC_E_UNSUPPORTED, "%s: ndim %d outside [1,%d]", who, ndim,
                  HEPMC_MAX_DIM);
    HEPMC_REQUIRE(K >= 1 && K <= HEPMC_MAX_DIVISIONS, HEPMC_E_UNSUPPORTED, "%s: divisions %d outside [1,%d]", who, K,
                  HEPMC_MAX_DIVISIONS);
    HEPMC_REQUIRE(n >= 0 && first_index >= 0, HEPMC_E_ARG, "%s: negative n/first_index", who);
    HEPMC_REQUIRE(chunk >= 256 && chunk % 256 == 0, HEPMC_E_ARG, "%s: chunk must be a positive multiple of 256", who);
    HEPMC_REQUIRE(!aligned || first_index % chunk == 0, HEPMC_E_ARG, "%s: first_index must be a multiple of chunk", who);
    return 0;
}

#define HEPMC_GENERIC_CASE(KIND)                                                                            \
    case KIND:                                                                                              \
        if (vegas) {                                                                                        \
            HEPMC_CUDA(cudaFuncSetAttribute(vegas_sample_kernel<KIND>,                                      \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            vegas_sample_kernel<KIND><<<blocks, threads, smem, st>>>(dens, a);                              \
        } else {                                                                                            \
            plain_sample_kernel<KIND, 0, false><<<blocks, 256, 0, st>>>(dens, a);                           \
        }                                                                                                   \
        break;

static int max_optin_smem() {
    static int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cached_dev = dev;
    }
    return cached;
}

// fast = native RNG: compile-time (kind, ndim, rng mode) instance when one exists (ndim <= 16), with or
// without per-point outputs; else the generic kernel (runtime ndim, replay tapes).
static int launch_sample(bool vegas, bool fast, const Dens& dens, const SampleArgs& a, unsigned blocks, int threads,
                         size_t smem, cudaStream_t st) {
    if (fast && vegas) {
        // the throughput kernel has its own shared-memory layout (256 threads); huge grids take the generic one
        const size_t fs = vegas_fast_smem_bytes(a.ndim, a.K, 256);
        const bool camel = dens.kind == HEPMC_CAMEL || dens.kind == HEPMC_UCAMEL;
        if (a.ndim > 16 || (int64_t)fs > (int64_t)max_optin_smem() || a.chunk % 256 != 0 ||
            (camel && !camel_uniform_means(dens)) || (dens.kind != HEPMC_NONE && a.idx_out)) fast = false;
        else smem = fs;
    }
    if (fast) {
        const bool out = a.x_out || a.f_out || a.w_out || a.idx_out;
        int e = -100;
        switch (dens.kind) {
            case HEPMC_CAMEL: e = launch_fast_camel(vegas, out, a.ndim, dens, a, blocks, threads, smem, st); break;
            case HEPMC_UCAMEL: e = launch_fast_ucamel(vegas, out, a.ndim, dens, a, blocks, threads, smem, st); break;
            case HEPMC_BANANA: e = launch_fast_banana(vegas, out, a.ndim, dens, a, blocks, threads, smem, st); break;
            case HEPMC_GAUSSIAN: e = launch_fast_gaussian(vegas, out, a.ndim, dens, a, blocks, threads, smem, st); break;
            case HEPMC_UNIFORM: e = launch_fast_uniform(vegas, out, a.ndim, dens, a, blocks, threads, smem, st); break;
            case HEPMC_NONE: e = launch_fast_none(vegas, out, a.ndim, dens, a, blocks, threads, smem, st); break;
            default: break;
        }
        if (e != -100) return e;
    }
    switch (dens.kind) {
        HEPMC_GENERIC_CASE(HEPMC_CAMEL)
        HEPMC_GENERIC_CASE(HEPMC_UCAMEL)
        HEPMC_GENERIC_CASE(HEPMC_BANANA)
        HEPMC_GENERIC_CASE(HEPMC_GAUSSIAN)
        HEPMC_GENERIC_CASE(HEPMC_UNIFORM)
        HEPMC_GENERIC_CASE(HEPMC_NONE)
        default:
            set_error("unknown density kind %d", dens.kind);
            return HEPMC_E_ARG;
    }
    return 0;
}

}  // namespace hepmc

using namespace hepmc;

extern "C" {

int64_t hepmc_default_chunk(int64_t n_total) {
    // a function of the iteration's TOTAL point count only (never of the GPU count)
    if (n_total >= ((int64_t)1 << 26)) return 32768;
    if (n_total >= ((int64_t)1 << 21)) return 8192;
    return 2048;
}

int hepmc_plain_mc_sample(const hepmc_density_t* dens, int64_t n, int64_t first_index, int64_t chunk,
                          uint64_t seed, uint32_t stream_id, int rng_mode, const double* u_replay_dev,
                          double* data_dev, double* f_dev, double* partials_dev, void* stream) {
    HEPMC_REQUIRE(dens != nullptr, HEPMC_E_ARG, "hepmc_plain_mc_sample: dens is NULL");
    int e = check_sample_common("hepmc_plain_mc_sample", dens->ndim, 1, n, first_index, chunk,
                                dens->kind != HEPMC_NONE);
    if (e) return e;
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_NONE, HEPMC_E_ARG, "bad density kind %d", dens->kind);
    HEPMC_REQUIRE(dens->kind == HEPMC_NONE || partials_dev, HEPMC_E_ARG, "hepmc_plain_mc_sample: partials_dev is NULL");
    HEPMC_REQUIRE(dens->kind != HEPMC_NONE || data_dev, HEPMC_E_ARG, "hepmc_plain_mc_sample: nothing to do");
    HEPMC_REQUIRE(rng_mode >= 0 && rng_mode < HEPMC_RNG_MODES, HEPMC_E_ARG, "hepmc_plain_mc_sample: unknown rng_mode %d", rng_mode);
    if (n == 0) return 0;
    SampleArgs a{};
    a.ndim = dens->ndim; a.K = 1; a.n = n; a.first_index = first_index; a.chunk = chunk;
    a.seed = seed; a.stream_id = stream_id; a.u_replay = u_replay_dev;
    a.rng_mode = rng_mode;
    a.keys = philox_expand(seed);
    a.x_out = data_dev; a.f_out = f_dev; a.stats_partials = partials_dev;
    const int64_t blocks = (n + chunk - 1) / chunk;
    const bool fast = !u_replay_dev;   // native draws: templated instance, with outputs if any pointer is given
    int e2 = launch_sample(false, fast, *dens, a, (unsigned)blocks, 256, 0, as_stream(stream));
    if (e2) return e2;
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_stats_partials(const double* values_dev, const double* weights_dev, double weight_scalar,
                         int64_t n, int64_t chunk, double* partials_dev, void* stream) {
    HEPMC_REQUIRE(values_dev && partials_dev && n >= 0, HEPMC_E_ARG, "hepmc_stats_partials: bad arguments");
    HEPMC_REQUIRE(chunk >= 256 && chunk % 256 == 0, HEPMC_E_ARG, "hepmc_stats_partials: bad chunk");
    if (n == 0) return 0;
    const int64_t blocks = (n + chunk - 1) / chunk;
    stats_partials_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(values_dev, weights_dev, weight_scalar, n,
                                                                           chunk, partials_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_reduce_stats(const double* rows_dev, const int64_t* row_begin_host, int n_out,
                       double* out_dev, void* stream) {
    RowRanges rr;
    int e = fill_ranges(rr, row_begin_host, n_out, "hepmc_reduce_stats");
    if (e) return e;
    HEPMC_REQUIRE(rows_dev && out_dev, HEPMC_E_ARG, "hepmc_reduce_stats: NULL buffer");
    reduce_stats_kernel<<<n_out, 256, 0, as_stream(stream)>>>(rows_dev, rr, out_dev, 3);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_reduce_hist(const double* rows_dev, const int64_t* row_begin_host, int n_out, int ncols,
                      double* out_dev, void* stream) {
    RowRanges rr;
    int e = fill_ranges(rr, row_begin_host, n_out, "hepmc_reduce_hist");
    if (e) return e;
    HEPMC_REQUIRE(rows_dev && out_dev && ncols >= 1, HEPMC_E_ARG, "hepmc_reduce_hist: bad arguments");
    dim3 grid((ncols + kHistCols - 1) / kHistCols, n_out);
    reduce_hist_kernel<<<grid, kHistLanes * kHistCols, 0, as_stream(stream)>>>(rows_dev, rr, ncols, out_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_vegas_sample(const hepmc_density_t* dens, const double* grid_dev, int ndim, int divisions,
                       int64_t n, int64_t first_index, int64_t chunk, uint64_t seed, uint32_t stream_id, int rng_mode,
                       const int64_t* idx_replay_dev, const double* u_replay_dev,
                       double* x_dev, double* f_dev, double* w_dev, int64_t* idx_dev,
                       double* stats_partials_dev, double* hist_partials_dev, void* stream) {
    HEPMC_REQUIRE(dens != nullptr && grid_dev != nullptr, HEPMC_E_ARG, "hepmc_vegas_sample: NULL dens/grid");
    int e = check_sample_common("hepmc_vegas_sample", ndim, divisions, n, first_index, chunk);
    if (e) return e;
    HEPMC_REQUIRE(dens->kind >= HEPMC_CAMEL && dens->kind <= HEPMC_NONE, HEPMC_E_ARG, "bad density kind %d", dens->kind);
    HEPMC_REQUIRE(dens->kind == HEPMC_NONE || dens->ndim == ndim, HEPMC_E_ARG,
                  "hepmc_vegas_sample: density ndim %d != grid ndim %d", dens->ndim, ndim);
    HEPMC_REQUIRE(rng_mode >= 0 && rng_mode < HEPMC_RNG_MODES, HEPMC_E_ARG, "hepmc_vegas_sample: unknown rng_mode %d", rng_mode);
    HEPMC_REQUIRE((idx_replay_dev == nullptr) == (u_replay_dev == nullptr), HEPMC_E_ARG,
                  "hepmc_vegas_sample: replay needs both idx and u tapes");
    const bool fused = dens->kind != HEPMC_NONE;
    HEPMC_REQUIRE(!fused || (stats_partials_dev && hist_partials_dev), HEPMC_E_ARG,
                  "hepmc_vegas_sample: partial buffers are NULL");
    HEPMC_REQUIRE(fused || (x_dev && w_dev && idx_dev), HEPMC_E_ARG,
                  "hepmc_vegas_sample: two-kernel path needs x, w and idx outputs");
    if (n == 0) return 0;
    size_t smem = 0;
    const int threads = pick_threads(ndim, divisions, smem);
    HEPMC_REQUIRE(threads > 0, HEPMC_E_UNSUPPORTED, "hepmc_vegas_sample: grid %d x %d does not fit shared memory", ndim,
                  divisions);
    HEPMC_REQUIRE(chunk % threads == 0, HEPMC_E_ARG, "chunk not a multiple of the block size");
    SampleArgs a{};
    a.grid = grid_dev; a.ndim = ndim; a.K = divisions; a.n = n; a.first_index = first_index; a.chunk = chunk;
    a.seed = seed; a.stream_id = stream_id; a.idx_replay = idx_replay_dev; a.u_replay = u_replay_dev;
    a.x_out = x_dev; a.f_out = f_dev; a.w_out = w_dev; a.idx_out = idx_dev;
    a.stats_partials = stats_partials_dev; a.hist_partials = hist_partials_dev;
    a.kpow = pow((double)divisions, (double)ndim);
    a.kpow_scaled = ldexp(a.kpow, 32 * (ndim > 8 ? ndim - 8 : ndim));
    a.rng_mode = rng_mode;
    a.keys = philox_expand(seed);
    set_camel_centre(a, *dens);
    const int64_t blocks = (n + chunk - 1) / chunk;
    const bool fast = !u_replay_dev;   // native draws: templated instance, with outputs if any pointer is given
    int e2 = launch_sample(true, fast, *dens, a, (unsigned)blocks, threads, smem, as_stream(stream));
    if (e2) return e2;
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_vegas_accumulate(const double* f_dev, const double* w_dev, const int64_t* idx_dev,
                           int ndim, int divisions, int64_t n, int64_t chunk,
                           double* stats_partials_dev, double* hist_partials_dev, void* stream) {
    int e = check_sample_common("hepmc_vegas_accumulate", ndim, divisions, n, 0, chunk);
    if (e) return e;
    HEPMC_REQUIRE(f_dev && w_dev && idx_dev && stats_partials_dev && hist_partials_dev, HEPMC_E_ARG,
                  "hepmc_vegas_accumulate: NULL buffer");
    if (n == 0) return 0;
    size_t smem = 0;
    const int threads = pick_threads(ndim, divisions, smem);
    HEPMC_REQUIRE(threads > 0, HEPMC_E_UNSUPPORTED, "hepmc_vegas_accumulate: %d divisions do not fit shared memory",
                  divisions);
    smem = vegas_smem_bytes(0, divisions, threads);
    HEPMC_CUDA(cudaFuncSetAttribute(vegas_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SampleArgs a{};
    a.ndim = ndim; a.K = divisions; a.n = n; a.chunk = chunk;
    a.f_in = f_dev; a.w_in = w_dev; a.idx_in = idx_dev;
    a.stats_partials = stats_partials_dev; a.hist_partials = hist_partials_dev;
    const int64_t blocks = (n + chunk - 1) / chunk;
    vegas_accumulate_kernel<<<(unsigned)blocks, threads, smem, as_stream(stream)>>>(a);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_vegas_update_grid(const double* slot_stats_dev, const double* slot_hist_dev, int n_slots,
                            double* grid_dev, int ndim, int divisions, double c, int j,
                            double* est_var_dev, void* stream) {
    HEPMC_REQUIRE(slot_stats_dev && slot_hist_dev && grid_dev && est_var_dev, HEPMC_E_ARG,
                  "hepmc_vegas_update_grid: NULL buffer");
    HEPMC_REQUIRE(n_slots >= 1 && ndim >= 1 && ndim <= HEPMC_MAX_DIM && divisions >= 1 && j >= 0, HEPMC_E_ARG,
                  "hepmc_vegas_update_grid: bad sizes");
    const size_t smem = (size_t)ndim * divisions * 8;
    HEPMC_CUDA(cudaFuncSetAttribute(vegas_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vegas_update_kernel<<<1, 256, smem, as_stream(stream)>>>(slot_stats_dev, 3, slot_hist_dev, (int64_t)ndim * divisions,
                                                             n_slots, grid_dev, ndim, divisions, c, j, est_var_dev, nullptr, 0,
                                                             nullptr);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int64_t hepmc_reduce_slots_scratch_bytes(int n_out, int ncols) {
    return (int64_t)n_out * kSlotGroups * (3 + ncols) * 8 + 256;   // group partials + per-slot arrival counters
}

int hepmc_reduce_slots(const double* stats_rows_dev, const double* hist_rows_dev, const int64_t* row_begin_host,
                       int n_out, int ncols, double* packed_dev, void* scratch_dev, void* stream) {
    RowRanges rr;
    int e = fill_ranges(rr, row_begin_host, n_out, "hepmc_reduce_slots");
    if (e) return e;
    HEPMC_REQUIRE(stats_rows_dev && hist_rows_dev && packed_dev && scratch_dev && ncols >= 1, HEPMC_E_ARG,
                  "hepmc_reduce_slots: bad arguments");
    const int ntile = (ncols + kRedCols - 1) / kRedCols;
    dim3 grid(ntile + 1, kSlotGroups, n_out);
    double* partials = reinterpret_cast<double*>(scratch_dev);
    unsigned int* counters = reinterpret_cast<unsigned int*>(partials + (size_t)n_out * kSlotGroups * (3 + ncols));
    SlotPeers none{};
    none.world = 1;
    reduce_slots_kernel<<<grid, kRedCols * kRedLanes, 0, as_stream(stream)>>>(stats_rows_dev, hist_rows_dev, rr, ncols,
                                                                              partials, counters, packed_dev, none);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int64_t hepmc_slot_exchange_bytes(int ncols) {
    return (int64_t)2 * HEPMC_N_SLOTS * (3 + ncols) * 8 + (int64_t)2 * HEPMC_N_SLOTS * 8;
}

int hepmc_reduce_slots_p2p(const double* stats_rows_dev, const double* hist_rows_dev, const int64_t* row_begin_host,
                           int n_out, int ncols, double* packed_dev, void* scratch_dev,
                           const int64_t* peer_buffers_host, int world, int first_slot, uint64_t epoch, void* stream) {
    RowRanges rr;
    int e = fill_ranges(rr, row_begin_host, n_out, "hepmc_reduce_slots_p2p");
    if (e) return e;
    HEPMC_REQUIRE(stats_rows_dev && hist_rows_dev && packed_dev && scratch_dev && ncols >= 1 && peer_buffers_host, HEPMC_E_ARG,
                  "hepmc_reduce_slots_p2p: bad arguments");
    HEPMC_REQUIRE(world >= 2 && world <= HEPMC_N_SLOTS && first_slot >= 0 && first_slot + n_out <= HEPMC_N_SLOTS && epoch > 0,
                  HEPMC_E_ARG, "hepmc_reduce_slots_p2p: world %d / first_slot %d / epoch out of range", world, first_slot);
    SlotPeers peers{};
    for (int r = 0; r < world; ++r) {
        HEPMC_REQUIRE(peer_buffers_host[r] != 0, HEPMC_E_ARG, "hepmc_reduce_slots_p2p: peer buffer %d is NULL", r);
        peers.base[r] = reinterpret_cast<double*>((uintptr_t)peer_buffers_host[r]);
    }
    peers.world = world; peers.slot0 = first_slot; peers.epoch = epoch;
    const int ntile = (ncols + kRedCols - 1) / kRedCols;
    dim3 grid(ntile + 1, kSlotGroups, n_out);
    double* partials = reinterpret_cast<double*>(scratch_dev);
    unsigned int* counters = reinterpret_cast<unsigned int*>(partials + (size_t)n_out * kSlotGroups * (3 + ncols));
    reduce_slots_kernel<<<grid, kRedCols * kRedLanes, 0, as_stream(stream)>>>(stats_rows_dev, hist_rows_dev, rr, ncols,
                                                                              partials, counters, packed_dev, peers);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_vegas_update_grid_p2p(const double* exchange_dev, int ncols, uint64_t epoch, double* grid_dev, int ndim,
                                int divisions, double c, int j, double* est_var_dev, int* timeout_flag_dev, void* stream) {
    HEPMC_REQUIRE(exchange_dev && grid_dev && est_var_dev && epoch > 0, HEPMC_E_ARG, "hepmc_vegas_update_grid_p2p: bad arguments");
    HEPMC_REQUIRE(ndim >= 1 && ndim <= HEPMC_MAX_DIM && divisions >= 1 && j >= 0 && ncols == ndim * divisions, HEPMC_E_ARG,
                  "hepmc_vegas_update_grid_p2p: bad sizes");
    const size_t smem = (size_t)ndim * divisions * 8;
    const int64_t row = 3 + (int64_t)ncols;
    const double* packed = exchange_dev + (size_t)(epoch & 1) * HEPMC_N_SLOTS * row;
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(exchange_dev + (size_t)2 * HEPMC_N_SLOTS * row) +
                                      (epoch & 1) * HEPMC_N_SLOTS;
    HEPMC_CUDA(cudaFuncSetAttribute(vegas_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vegas_update_kernel<<<1, 256, smem, as_stream(stream)>>>(packed, row, packed + 3, row, HEPMC_N_SLOTS, grid_dev, ndim,
                                                             divisions, c, j, est_var_dev, flags, epoch, timeout_flag_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

int hepmc_vegas_update_grid_packed(const double* packed_dev, int n_slots, double* grid_dev, int ndim, int divisions,
                                   double c, int j, double* est_var_dev, void* stream) {
    HEPMC_REQUIRE(packed_dev && grid_dev && est_var_dev, HEPMC_E_ARG, "hepmc_vegas_update_grid_packed: NULL buffer");
    HEPMC_REQUIRE(n_slots >= 1 && ndim >= 1 && ndim <= HEPMC_MAX_DIM && divisions >= 1 && j >= 0, HEPMC_E_ARG,
                  "hepmc_vegas_update_grid_packed: bad sizes");
    const size_t smem = (size_t)ndim * divisions * 8;
    const int64_t row = 3 + (int64_t)ndim * divisions;
    HEPMC_CUDA(cudaFuncSetAttribute(vegas_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vegas_update_kernel<<<1, 256, smem, as_stream(stream)>>>(packed_dev, row, packed_dev + 3, row, n_slots, grid_dev, ndim,
                                                             divisions, c, j, est_var_dev, nullptr, 0, nullptr);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

static void slot_ranges(int64_t n_chunks, int64_t* begin /* HEPMC_N_SLOTS+1 */) {
    for (int s = 0; s <= HEPMC_N_SLOTS; ++s) begin[s] = (n_chunks * s) / HEPMC_N_SLOTS;
}

int64_t hepmc_vegas_workspace_bytes(int ndim, int divisions, int64_t n) {
    const int64_t chunk = hepmc_default_chunk(n);
    const int64_t nc = (n + chunk - 1) / chunk;
    const int64_t cols = (int64_t)ndim * divisions;
    return 8 * (nc * 3 + nc * cols + HEPMC_N_SLOTS * (3 + cols)) + hepmc_reduce_slots_scratch_bytes(HEPMC_N_SLOTS, (int)cols) + 256;
}

int hepmc_vegas_integrate(const hepmc_density_t* dens, double* grid_dev, int ndim, int divisions,
                          int64_t n, int iterations, double c, uint64_t seed, uint32_t stream_id0, int rng_mode,
                          double* est_var_dev, void* workspace_dev, int64_t workspace_bytes,
                          void* stream) {
    HEPMC_REQUIRE(dens && grid_dev && est_var_dev && workspace_dev, HEPMC_E_ARG, "hepmc_vegas_integrate: NULL buffer");
    HEPMC_REQUIRE(dens->kind != HEPMC_NONE, HEPMC_E_ARG, "hepmc_vegas_integrate needs a built-in integrand");
    HEPMC_REQUIRE(n > 0 && iterations >= 1, HEPMC_E_ARG, "hepmc_vegas_integrate: n and iterations must be positive");
    HEPMC_REQUIRE(workspace_bytes >= hepmc_vegas_workspace_bytes(ndim, divisions, n), HEPMC_E_ARG,
                  "hepmc_vegas_integrate: workspace too small");
    const int64_t chunk = hepmc_default_chunk(n);
    const int64_t nc = (n + chunk - 1) / chunk;
    const int64_t cols = (int64_t)ndim * divisions;
    double* ws = reinterpret_cast<double*>(workspace_dev);
    double* stats_p = ws;
    double* hist_p = stats_p + nc * 3;
    double* packed = hist_p + nc * cols;
    double* scratch = packed + HEPMC_N_SLOTS * (3 + cols);
    int64_t begin[HEPMC_N_SLOTS + 1];
    slot_ranges(nc, begin);
    HEPMC_CUDA(cudaMemsetAsync(scratch + (size_t)HEPMC_N_SLOTS * kSlotGroups * (3 + cols), 0, 256, as_stream(stream)));
    // small problems (at most kSlotGroups chunks per slot): reduction and adaptation in one single-CTA launch
    bool small = true;
    for (int s = 0; s < HEPMC_N_SLOTS; ++s) small = small && (begin[s + 1] - begin[s] <= kSlotGroups);
    RowRanges rr;
    if (small) {
        int e = fill_ranges(rr, begin, HEPMC_N_SLOTS, "hepmc_vegas_integrate");
        if (e) return e;
        HEPMC_CUDA(cudaFuncSetAttribute(vegas_small_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cols * 8)));
    }
    for (int j = 0; j < iterations; ++j) {
        int e = hepmc_vegas_sample(dens, grid_dev, ndim, divisions, n, 0, chunk, seed, stream_id0 + (uint32_t)j, rng_mode,
                                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stats_p, hist_p, stream);
        if (e) return e;
        if (small) {
            vegas_small_step_kernel<<<1, 256, (size_t)cols * 8, as_stream(stream)>>>(stats_p, hist_p, rr, (int)cols, packed, grid_dev,
                                                                                   ndim, divisions, c, j, est_var_dev);
            HEPMC_LAUNCH_CHECK();
            continue;
        }
        e = hepmc_reduce_slots(stats_p, hist_p, begin, HEPMC_N_SLOTS, (int)cols, packed, scratch, stream);
        if (e) return e;
        e = hepmc_vegas_update_grid_packed(packed, HEPMC_N_SLOTS, grid_dev, ndim, divisions, c, j, est_var_dev, stream);
        if (e) return e;
    }
    return 0;
}

int hepmc_grid_pdf(const double* grid_dev, int ndim, int divisions, const double* xs_dev,
                   int64_t n, double* out_dev, void* stream) {
    HEPMC_REQUIRE(grid_dev && ndim >= 1 && ndim <= HEPMC_MAX_DIM && divisions >= 1 && n >= 0, HEPMC_E_ARG,
                  "hepmc_grid_pdf: bad arguments");
    if (n == 0) return 0;
    HEPMC_REQUIRE(xs_dev && out_dev, HEPMC_E_ARG, "hepmc_grid_pdf: NULL buffer");
    const size_t smem = (size_t)ndim * (2 * divisions + 1) * 8;
    HEPMC_CUDA(cudaFuncSetAttribute(grid_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    grid_pdf_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(grid_dev, ndim, divisions,
                                                                         pow((double)divisions, (double)ndim), xs_dev, n,
                                                                         out_dev);
    HEPMC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
