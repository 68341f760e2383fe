// Detection tail: box decode + clip, per-class score threshold -> sort -> greedy NMS ->
// cross-class top-k + pad.  Integer / index work, bit-exact w.r.t. oracle/tail.py.
// Compiled with -fmad=false: every float op is individually rounded like the reference's
// separate TF ops (RegressBoxes.py:150-162) and TF's CPU NMS kernel.
//
// Replaces (reference file:line):
//   RegressBoxes.py:126-164, ClipBoxes.py:9-24               -> boxes_kernel
//   FilterDetections.py:9 (tf.where over score>thr)          -> scan_scores_kernel + fill_keys_kernel (scan_rows_kernel)
//   FilterDetections.py:16-21 (tf.image.non_max_suppression) -> sort_nms_{small,large}_kernel
//   FilterDetections.py:85-111 (concat, top_k, gather, pad)  -> merge_topk_kernel
#include "common.cuh"

namespace effdet {

typedef unsigned long long u64;

// ------------------------------------------------------------------ decode + clip
template <bool REGRESS, bool CLIP>
__global__ void __launch_bounds__(256)
boxes_kernel(const float4 *__restrict__ anchors, int anchors_per_image,
             const float4 *__restrict__ in, float4 mean, float4 stdv, size_t N, size_t total,
             float hi_x, float hi_y, float4 *__restrict__ out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        float4 v = in[i];
        if (REGRESS) {
            float4 a = anchors[anchors_per_image ? i : (i % N)];
            float w = __fsub_rn(a.z, a.x), h = __fsub_rn(a.w, a.y);
            v.x = __fadd_rn(a.x, __fmul_rn(__fadd_rn(__fmul_rn(v.x, stdv.x), mean.x), w));
            v.y = __fadd_rn(a.y, __fmul_rn(__fadd_rn(__fmul_rn(v.y, stdv.y), mean.y), h));
            v.z = __fadd_rn(a.z, __fmul_rn(__fadd_rn(__fmul_rn(v.z, stdv.z), mean.z), w));
            v.w = __fadd_rn(a.w, __fmul_rn(__fadd_rn(__fmul_rn(v.w, stdv.w), mean.w), h));
        }
        if (CLIP) {   // tf.clip_by_value = max(min(x, hi), lo)
            v.x = fmaxf(fminf(v.x, hi_x), 0.f);
            v.y = fmaxf(fminf(v.y, hi_y), 0.f);
            v.z = fmaxf(fminf(v.z, hi_x), 0.f);
            v.w = fmaxf(fminf(v.w, hi_y), 0.f);
        }
        out[i] = v;
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool REGRESS, bool CLIP>
static int launch_boxes(const float *anchors, int per_image, const float *in, const float *mean,
                        const float *stdv, int B, size_t N, float H, float W, float *out,
                        void *stream) {
    EFFDET_REQUIRE(B >= 0, "negative batch");
    size_t total = (size_t)B * N;
    if (total == 0) return EFFDET_OK;
    EFFDET_REQUIRE(in && out, "null pointer");
    EFFDET_REQUIRE(aligned16(in) && aligned16(out), "boxes must be 16-byte aligned");
    float4 m = make_float4(0, 0, 0, 0), s = make_float4(1, 1, 1, 1);
    if (REGRESS) {
        EFFDET_REQUIRE(anchors && mean && stdv, "null pointer");
        EFFDET_REQUIRE(aligned16(anchors), "anchors must be 16-byte aligned");
        m = make_float4(mean[0], mean[1], mean[2], mean[3]);
        s = make_float4(stdv[0], stdv[1], stdv[2], stdv[3]);
    }
    unsigned blocks = cdiv(total, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    boxes_kernel<REGRESS, CLIP><<<blocks, 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4 *>(anchors), per_image, reinterpret_cast<const float4 *>(in),
        m, s, N, total, W - 1.f, H - 1.f, reinterpret_cast<float4 *>(out));
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}

// ------------------------------------------------------------------ candidate keys
// key = (~orderable(score)) << 32 | anchor_index : ascending u64 order == (score desc, index asc)
__device__ __forceinline__ uint32_t orderable(float s) {
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorderable(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ u64 make_key(float s, uint32_t idx) {
    return ((u64)(~orderable(s)) << 32) | idx;
}
__device__ __forceinline__ float key_score(u64 k) { return unorderable(~(uint32_t)(k >> 32)); }

constexpr int kScanThreads = 256;
constexpr int kScanElems = 4096;   // floats per block

// Count pass of the class-specific scan.  One block = kScanElems consecutive floats of image blockIdx.y (coalesced);
// class of element e is e % C.  Writes per-(image,class) counts and ONE BIT per passing score (score > thr) into the
// zero-filled `mask` (cudaMemsetAsync by the launcher, 1/32 of the scores' bytes; a rare global atomicOr).  The loop
// is the plain round-1 loop (load, compare, rare atomics) and runs at the DRAM rate (6.5 TB/s); collecting the bits
// per block in shared memory slowed it from 0.39 to ~0.6 ms, by warp ballot or from tied load batches more still.
// The fill pass (fill_keys_kernel) then never touches the (B, N, C) scores again except for the passing ones: the
// second full pass over them was 0.69 ms of the 1.2 ms tail at D2 / batch 64 / 90 classes (profiles/r2_tail_ncu.txt).
__global__ void __launch_bounds__(kScanThreads)
scan_scores_kernel(const float *__restrict__ cls, uint32_t NC, int C, float thr, uint32_t *__restrict__ counts,
                   uint32_t *__restrict__ mask, uint32_t mask_words) {
    extern __shared__ uint32_t sh[];
    uint32_t *hist = sh;
    const int b = blockIdx.y, tid = threadIdx.x;
    const float *p = cls + (size_t)b * NC;
    uint32_t *m = mask + (size_t)b * mask_words;
    const uint32_t e0 = blockIdx.x * (uint32_t)kScanElems;
    const uint32_t e1 = min(e0 + (uint32_t)kScanElems, NC);
    for (int c = tid; c < C; c += kScanThreads) hist[c] = 0;
    __syncthreads();
    const uint32_t step = kScanThreads % C;
    uint32_t c = (e0 + tid) % C;
    for (uint32_t e = e0 + tid; e < e1; e += kScanThreads) {
        const float sc = p[e];
        if (sc > thr) {
            atomicAdd(&hist[c], 1u);
            atomicOr(&m[e >> 5], 1u << (e & 31u));
        }
        c += step; if (c >= (uint32_t)C) c -= C;
    }
    __syncthreads();
    for (int k = tid; k < C; k += kScanThreads)
        if (hist[k]) atomicAdd(&counts[(size_t)b * C + k], hist[k]);
}

// Fill pass of the class-specific scan, reading only the mask: a block owns kFillWords mask words (32768 scores),
// a thread four of them; the work is one coalesced load per 32 scores plus a loop over the few set bits.  (The first
// mask-reading version kept one block per 4096 scores and rebuilt per-thread bit vectors with 16 shuffles: 155 000
// blocks at the block-launch rate, issue-bound, 245 us for 80 MB at D2 / batch 64; this one: 19 us.)
constexpr int kFillWords = 1024;
__global__ void __launch_bounds__(kScanThreads)
fill_keys_kernel(const float *__restrict__ cls, uint32_t NC, int C, const uint32_t *__restrict__ offsets,
                 uint32_t *__restrict__ cursor, u64 *__restrict__ keys, const int32_t *__restrict__ status,
                 const uint32_t *__restrict__ mask, uint32_t mask_words) {
    extern __shared__ uint32_t sh[];
    uint32_t *hist = sh, *base = sh + C;
    if (status[0]) return;
    const int b = blockIdx.y, tid = threadIdx.x;
    const float *p = cls + (size_t)b * NC;
    const uint32_t *m = mask + (size_t)b * mask_words;
    const uint32_t w0 = blockIdx.x * (uint32_t)kFillWords;
    constexpr int kPerT = kFillWords / kScanThreads;
    uint32_t wd[kPerT];
    bool any = false;
#pragma unroll
    for (int j = 0; j < kPerT; ++j) {
        const uint32_t wi = w0 + (uint32_t)j * kScanThreads + tid;
        wd[j] = wi < mask_words ? m[wi] : 0u;
        any |= wd[j] != 0u;
    }
    for (int c = tid; c < C; c += kScanThreads) hist[c] = 0;
    if (!__syncthreads_or(any)) return;                  // no candidate in these 32768 scores
#pragma unroll
    for (int j = 0; j < kPerT; ++j)
        for (uint32_t t = wd[j]; t; t &= t - 1) {
            const uint32_t e = (w0 + (uint32_t)j * kScanThreads + tid) * 32u + (uint32_t)(__ffs(t) - 1);
            atomicAdd(&hist[e % (uint32_t)C], 1u);
        }
    __syncthreads();
    for (int k = tid; k < C; k += kScanThreads) {
        const uint32_t h = hist[k];
        base[k] = h ? offsets[(size_t)b * C + k] + atomicAdd(&cursor[(size_t)b * C + k], h) : 0u;
        hist[k] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPerT; ++j)
        for (uint32_t t = wd[j]; t; t &= t - 1) {
            const uint32_t e = (w0 + (uint32_t)j * kScanThreads + tid) * 32u + (uint32_t)(__ffs(t) - 1);
            const uint32_t cj = e % (uint32_t)C;
            const float sc = p[e];
            const uint32_t r = atomicAdd(&hist[cj], 1u);
            keys[base[cj] + r] = make_key(sc, e / (uint32_t)C);
        }
}

// class-agnostic scan (class_specific_filter=False, FilterDetections.py:86-92): score = row max.
template <bool FILL>
__global__ void __launch_bounds__(kScanThreads)
scan_rows_kernel(const float *__restrict__ cls, uint32_t N, int C, float thr,
                 uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets,
                 uint32_t *__restrict__ cursor, u64 *__restrict__ keys,
                 const int32_t *__restrict__ status) {
    if (FILL && status[0]) return;
    const int b = blockIdx.y;
    const uint32_t n = blockIdx.x * kScanThreads + threadIdx.x;
    bool pass = false;
    float mx = 0.f;
    if (n < N) {
        const float *row = cls + ((size_t)b * N + n) * C;
        mx = row[0];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
        pass = mx > thr;
    }
    unsigned ballot = __ballot_sync(0xffffffffu, pass);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(ballot) - 1;
    uint32_t pos = 0;
    if (lane == leader) pos = atomicAdd(FILL ? &cursor[b] : &counts[b], __popc(ballot));
    if (!FILL) return;
    pos = __shfl_sync(0xffffffffu, pos, leader);
    if (pass) keys[offsets[b] + pos + __popc(ballot & ((1u << lane) - 1))] = make_key(mx, n);
}

// exclusive scan of the per-segment counts (single block) + overflow check
__global__ void __launch_bounds__(1024)
offsets_kernel(const uint32_t *__restrict__ counts, uint32_t nseg, uint32_t *__restrict__ offsets,
               uint32_t *__restrict__ cursor, u64 capacity, int32_t *__restrict__ status) {
    __shared__ u64 warp_sums[32];
    __shared__ u64 carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nseg; base += 1024) {
        uint32_t i = base + tid;
        u64 v = i < nseg ? counts[i] : 0, x = v;
        for (int d = 1; d < 32; d <<= 1) {
            u64 y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            u64 w = warp_sums[lane], ws = w;
            for (int d = 1; d < 32; d <<= 1) {
                u64 y = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += y;
            }
            warp_sums[lane] = ws - w;    // exclusive
        }
        __syncthreads();
        u64 excl = carry + warp_sums[warp] + (x - v);
        if (i < nseg) {
            offsets[i] = (uint32_t)(excl > 0xffffffffull ? 0xffffffffull : excl);
            cursor[i] = 0;
        }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        offsets[nseg] = (uint32_t)(carry > 0xffffffffull ? 0xffffffffull : carry);
        status[0] = carry > capacity ? 1 : 0;
        status[1] = (int32_t)(carry > 0x7fffffffull ? 0x7fffffffull : carry);
    }
}

// ------------------------------------------------------------------ sort + NMS
// All-ascending bitonic network ("flip" first step per merge level, then half-cleaners):
// every comparator leaves the minimum at the lower index, so positions >= len behave as
// virtual +inf padding and comparators touching them can simply be skipped.
template <int T>
__device__ __forceinline__ void bitonic_sort_inplace(u64 *k, uint32_t len) {
    uint32_t P = 1;
    while (P < len) P <<= 1;
    for (uint32_t size = 2; size <= P; size <<= 1) {
        const uint32_t half = size >> 1;
        for (uint32_t i = threadIdx.x; i < (P >> 1); i += T) {
            uint32_t blk = i / half, r = i % half;
            uint32_t lo = blk * size + r, hi = blk * size + size - 1 - r;
            if (hi < len) {
                u64 a = k[lo], b = k[hi];
                if (a > b) { k[lo] = b; k[hi] = a; }
            }
        }
        __syncthreads();
        for (uint32_t j = half >> 1; j >= 1; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < (P >> 1); i += T) {
                uint32_t lo = 2 * j * (i / j) + (i % j), hi = lo + j;
                if (hi < len) {
                    u64 a = k[lo], b = k[hi];
                    if (a > b) { k[lo] = b; k[hi] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// TF NonMaxSuppressionV3 IoU (float32, corner order normalised, zero-area => 0, no "+1")
__device__ __forceinline__ bool iou_gt(float4 a, float4 b, float thr) {
    float a0 = fminf(a.x, a.z), a2 = fmaxf(a.x, a.z), a1 = fminf(a.y, a.w), a3 = fmaxf(a.y, a.w);
    float b0 = fminf(b.x, b.z), b2 = fmaxf(b.x, b.z), b1 = fminf(b.y, b.w), b3 = fmaxf(b.y, b.w);
    float area_a = __fmul_rn(__fsub_rn(a2, a0), __fsub_rn(a3, a1));
    float area_b = __fmul_rn(__fsub_rn(b2, b0), __fsub_rn(b3, b1));
    if (area_a <= 0.f || area_b <= 0.f) return false;
    float i0 = fmaxf(a0, b0), i1 = fmaxf(a1, b1), i2 = fminf(a2, b2), i3 = fminf(a3, b3);
    float inter = __fmul_rn(fmaxf(__fsub_rn(i2, i0), 0.f), fmaxf(__fsub_rn(i3, i1), 0.f));
    float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return iou > thr;
}

constexpr int kNmsTile = 64;

// Greedy NMS over `len` keys already sorted (score desc, index asc) at `sorted` (shared or
// global).  Kept keys are written, in selection order, to out[0..nsel): out may alias the
// global segment the keys came from (write position never passes the read position, and a
// tile's keys are staged in shared memory before anything is written).
// T threads, T % 64 == 0.  sel: shared float4[max_det].
template <int T>
__device__ uint32_t greedy_nms(const u64 *sorted, uint32_t len, const float4 *__restrict__ boxes_b,
                               float thr, uint32_t max_det, float4 *sel, u64 *out) {
    __shared__ float4 tb[kNmsTile];
    __shared__ u64 tk[kNmsTile];
    __shared__ u64 tmask[kNmsTile];
    __shared__ u64 dead, sel_mask;
    __shared__ uint32_t nsel_sh;
    const int tid = threadIdx.x, i = tid & (kNmsTile - 1), q = tid / kNmsTile;
    constexpr int Q = T / kNmsTile;
    if (tid == 0) nsel_sh = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < len; t0 += kNmsTile) {
        const uint32_t nsel = nsel_sh;
        if (nsel >= max_det) break;
        const uint32_t tn = min((uint32_t)kNmsTile, len - t0);
        if (tid < kNmsTile) {
            tmask[tid] = 0;
            if ((uint32_t)tid < tn) {
                u64 k = sorted[t0 + tid];
                tk[tid] = k;
                tb[tid] = boxes_b[(uint32_t)k];
            }
        }
        if (tid == 0) dead = 0;
        __syncthreads();
        if ((uint32_t)i < tn) {
            const float4 me = tb[i];
            bool d = false;
            for (uint32_t j = q; j < nsel && !d; j += Q) d = iou_gt(me, sel[j], thr);
            if (d) atomicOr(&dead, 1ull << i);
            u64 bits = 0;
            constexpr int PER = kNmsTile / Q;
            for (int j = q * PER; j < (q + 1) * PER && j < i; ++j)
                if (iou_gt(me, tb[j], thr)) bits |= 1ull << j;
            if (bits) atomicOr(&tmask[i], bits);
        }
        __syncthreads();
        // The greedy pass over the tile is sequential only in the `selected` mask: thread 0 runs it as a register
        // loop (the tmask loads do not depend on it and are issued ahead by the unrolling), then 64 threads scatter
        // the selected candidates by popcount rank.  (Round 1 also did the stores inside the one-thread loop: with
        // ~18 tiles per segment that loop was 37 % of the kernel's stall samples, 255 threads waiting at the barrier.)
        // Candidates ranked past max_det are dropped; they are the last of the tile and the loop ends after it.
        if (tid == 0) {
            u64 selected = 0;
            const u64 dd = dead;
#pragma unroll 8
            for (uint32_t c = 0; c < tn; ++c) {
                const u64 mk = tmask[c];
                if (((dd >> c) & 1ull) == 0 && (mk & selected) == 0) selected |= 1ull << c;
            }
            sel_mask = selected;
        }
        __syncthreads();
        {
            const u64 selected = sel_mask;
            if (tid < kNmsTile && ((selected >> tid) & 1ull)) {
                const uint32_t pos = nsel + (uint32_t)__popcll(selected & ((1ull << tid) - 1ull));
                if (pos < max_det) { sel[pos] = tb[tid]; out[pos] = tk[tid]; }
            }
            if (tid == 0) nsel_sh = min(nsel + (uint32_t)__popcll(selected), max_det);
        }
        __syncthreads();
    }
    return nsel_sh;
}

constexpr int kSmallCap = 2048;
constexpr int kSmallThreads = 256;
constexpr int kLargeThreads = 1024;

// segments with 1..kSmallCap candidates: sort in shared memory
__global__ void __launch_bounds__(kSmallThreads)
sort_nms_small_kernel(const float4 *__restrict__ boxes, size_t N, u64 *__restrict__ keys,
                      const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ counts,
                      uint32_t *__restrict__ kept, int S, float iou_thr, uint32_t max_det,
                      int do_nms, const int32_t *__restrict__ status) {
    __shared__ u64 sk[kSmallCap];
    extern __shared__ float4 sel[];
    if (status[0]) return;
    const uint32_t seg = blockIdx.x, len = counts[seg];
    if (len == 0) { if (threadIdx.x == 0) kept[seg] = 0; return; }
    if (len > kSmallCap) return;
    u64 *gk = keys + offsets[seg];
    for (uint32_t i = threadIdx.x; i < len; i += kSmallThreads) sk[i] = gk[i];
    __syncthreads();
    bitonic_sort_inplace<kSmallThreads>(sk, len);
    if (!do_nms) {
        for (uint32_t i = threadIdx.x; i < len; i += kSmallThreads) gk[i] = sk[i];
        if (threadIdx.x == 0) kept[seg] = len;
        return;
    }
    uint32_t n = greedy_nms<kSmallThreads>(sk, len, boxes + (size_t)(seg / S) * N, iou_thr, max_det,
                                           sel, gk);
    if (threadIdx.x == 0) kept[seg] = n;
}

// segments with more than kSmallCap candidates: sort in place in global memory
__global__ void __launch_bounds__(kLargeThreads)
sort_nms_large_kernel(const float4 *__restrict__ boxes, size_t N, u64 *__restrict__ keys,
                      const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ counts,
                      uint32_t *__restrict__ kept, int S, float iou_thr, uint32_t max_det,
                      int do_nms, const int32_t *__restrict__ status) {
    extern __shared__ float4 sel[];
    if (status[0]) return;
    const uint32_t seg = blockIdx.x, len = counts[seg];
    if (len <= kSmallCap) return;
    u64 *gk = keys + offsets[seg];
    bitonic_sort_inplace<kLargeThreads>(gk, len);
    if (!do_nms) { if (threadIdx.x == 0) kept[seg] = len; return; }
    uint32_t n = greedy_nms<kLargeThreads>(gk, len, boxes + (size_t)(seg / S) * N, iou_thr, max_det,
                                           sel, gk);
    if (threadIdx.x == 0) kept[seg] = n;
}

// ------------------------------------------------------------------ cross-class top-k + pad
// One block per image: S-way merge of the per-segment kept lists (each already score-desc).
// tf.nn.top_k tie rule = lower position in the class-major concatenation = lower class first.
// Three phases.  (1) all threads stage the score words (high halves of the keys) of the first min(kept, max_det)
// entries of every list in shared memory; (2) warp 0 merges on those words alone -- a lane owns lists lane,
// lane + 32, ..., keeps their current heads in registers, the warp takes the minimum by shuffles, the owner advances
// in shared memory -- and records the picks (list, position); (3) all threads gather boxes / scores / labels of the
// picks.  The round-1 kernel did all of it in the merge loop of one warp per image: two dependent global loads per
// pick (the next key of the list that won, then its box), ~0.9 us x 300 picks = 270 us whatever the batch.
constexpr int kMergeThreads = 128;
constexpr int kMergeMaxLists = 256;          // lists per image on the staged path (NQ = 1, 3 or 8 lists per lane)
template <int NQ>
__global__ void __launch_bounds__(kMergeThreads)
merge_topk_kernel(const float4 *__restrict__ boxes, const float *__restrict__ cls, int B, size_t N,
                  int C, int S, const u64 *__restrict__ keys, const uint32_t *__restrict__ offsets,
                  const uint32_t *__restrict__ kept, int max_det, float4 *__restrict__ out_boxes,
                  float *__restrict__ out_scores, int32_t *__restrict__ out_labels,
                  int32_t *__restrict__ out_indices, const int32_t *__restrict__ status) {
    extern __shared__ uint32_t msh[];
    uint32_t *words = msh;                              // [S][max_det] score words of the staged entries
    uint32_t *picks = msh + (size_t)S * max_det;        // [max_det] (list << 16) | position
    __shared__ int produced_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    int produced = 0;
    if (!status[0]) {
        const uint32_t *off = offsets + (size_t)b * S, *kp = kept + (size_t)b * S;
        // ---- (1) stage
        for (int s = warp; s < S; s += kMergeThreads / 32) {
            const uint32_t n = min(kp[s], (uint32_t)max_det);
            const u64 *src = keys + off[s];
            for (uint32_t i = lane; i < n; i += 32) words[(size_t)s * max_det + i] = (uint32_t)(src[i] >> 32);
        }
        __syncthreads();
        // ---- (2) merge (warp 0).  Per pick: the lane's best score word over its lists, one REDUX for the warp's
        // best word, one more for the lowest class among the lanes that hold it (ties: lower class first), then the
        // owner advances that list in shared memory.  (With 64-bit keys and five shuffle rounds a pick cost ~660
        // cycles: 111 us per 300 detections.)
        if (warp == 0) {
            uint32_t hw[NQ];                            // current score word of list lane + 32 q, 0xffffffff = exhausted
            uint32_t pos[NQ], cnt[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int sidx = lane + 32 * q;
                pos[q] = 0;
                cnt[q] = sidx < S ? min(kp[sidx], (uint32_t)max_det) : 0u;
                hw[q] = cnt[q] ? words[(size_t)sidx * max_det] : 0xffffffffu;
            }
            for (; produced < max_det; ++produced) {
                uint32_t mw = hw[0];
#pragma unroll
                for (int q = 1; q < NQ; ++q) mw = min(mw, hw[q]);
                const uint32_t best = __reduce_min_sync(0xffffffffu, mw);
                if (best == 0xffffffffu) break;
                uint32_t myc = 0xffffffffu;             // lowest class of this lane whose head has the best word
#pragma unroll
                for (int q = NQ - 1; q >= 0; --q)
                    if (hw[q] == best) myc = (uint32_t)(lane + 32 * q);
                const uint32_t sidx = NQ == 1 ? (uint32_t)(__ffs(__ballot_sync(0xffffffffu, myc != 0xffffffffu)) - 1)
                                              : __reduce_min_sync(0xffffffffu, myc);
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    if (sidx == (uint32_t)(lane + 32 * q)) {
                        picks[produced] = (sidx << 16) | pos[q];
                        ++pos[q];
                        hw[q] = pos[q] < cnt[q] ? words[(size_t)sidx * max_det + pos[q]] : 0xffffffffu;
                    }
            }
            if (lane == 0) produced_sh = produced;
        }
        __syncthreads();
        produced = produced_sh;
        // ---- (3) gather
        for (int j = tid; j < produced; j += kMergeThreads) {
            const uint32_t pk = picks[j];
            const int sidx = (int)(pk >> 16);
            const u64 k = keys[off[sidx] + (pk & 0xffffu)];
            const uint32_t idx = (uint32_t)k;
            int label = sidx;
            if (S != C) {    // class-agnostic: label = first argmax of the row
                const float *row = cls + ((size_t)b * N + idx) * C;
                float mx = row[0]; label = 0;
                for (int c = 1; c < C; ++c) if (row[c] > mx) { mx = row[c]; label = c; }
            }
            const size_t o = (size_t)b * max_det + j;
            out_boxes[o] = boxes[(size_t)b * N + idx];
            out_scores[o] = key_score(k);
            out_labels[o] = label;
            if (out_indices) out_indices[o] = (int32_t)idx;
        }
    }
    for (int j = produced + tid; j < max_det; j += kMergeThreads) {
        size_t o = (size_t)b * max_det + j;
        out_boxes[o] = make_float4(-1.f, -1.f, -1.f, -1.f);
        out_scores[o] = -1.f;
        out_labels[o] = -1;
        if (out_indices) out_indices[o] = -1;
    }
}

// Fallback of the cross-class merge for shapes the staged kernel above does not take (very large max_detections
// without NMS, more than 256 lists): the round-1 kernel.
// One warp per image: S-way merge of the per-segment kept lists (each already score-desc).
// tf.nn.top_k tie rule = lower position in the class-major concatenation = lower class first.
constexpr int kMergeWarps = 4;
__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_warp_kernel(const float4 *__restrict__ boxes, const float *__restrict__ cls, int B, size_t N,
                  int C, int S, const u64 *__restrict__ keys, const uint32_t *__restrict__ offsets,
                  const uint32_t *__restrict__ kept, int max_det, float4 *__restrict__ out_boxes,
                  float *__restrict__ out_scores, int32_t *__restrict__ out_labels,
                  int32_t *__restrict__ out_indices, const int32_t *__restrict__ status) {
    extern __shared__ uint32_t heads_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kMergeWarps + warp;
    if (b >= B) return;
    uint32_t *head = heads_all + (size_t)warp * S;
    int produced = 0;
    if (!status[0]) {
        const uint32_t *off = offsets + (size_t)b * S, *kp = kept + (size_t)b * S;
        for (int s = lane; s < S; s += 32) head[s] = 0;
        __syncwarp();
        auto rescan = [&]() -> u64 {
            u64 best = ~0ull;
            for (int s = lane; s < S; s += 32) {
                uint32_t h = head[s];
                if (h < kp[s]) {
                    u64 m = (keys[off[s] + h] & 0xffffffff00000000ull) | (uint32_t)s;
                    best = m < best ? m : best;
                }
            }
            return best;
        };
        u64 mine = rescan();
        for (; produced < max_det; ++produced) {
            u64 m = mine;
            for (int d = 16; d >= 1; d >>= 1) {
                u64 o = __shfl_xor_sync(0xffffffffu, m, d);
                m = o < m ? o : m;
            }
            if (m == ~0ull) break;
            const int s = (int)(uint32_t)m;
            if ((s & 31) == lane) {
                uint32_t h = head[s];
                u64 k = keys[off[s] + h];
                uint32_t idx = (uint32_t)k;
                int label = s;
                if (S != C) {    // class-agnostic: label = first argmax of the row
                    const float *row = cls + ((size_t)b * N + idx) * C;
                    float mx = row[0]; label = 0;
                    for (int c = 1; c < C; ++c) if (row[c] > mx) { mx = row[c]; label = c; }
                }
                size_t o = (size_t)b * max_det + produced;
                out_boxes[o] = boxes[(size_t)b * N + idx];
                out_scores[o] = key_score(k);
                out_labels[o] = label;
                if (out_indices) out_indices[o] = (int32_t)idx;
                head[s] = h + 1;
                mine = rescan();
            }
        }
    }
    for (int j = produced + lane; j < max_det; j += 32) {
        size_t o = (size_t)b * max_det + j;
        out_boxes[o] = make_float4(-1.f, -1.f, -1.f, -1.f);
        out_scores[o] = -1.f;
        out_labels[o] = -1;
        if (out_indices) out_indices[o] = -1;
    }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct FdLayout {
    size_t counts, offsets, cursor, kept, keys, mask, mask_words, total;
    FdLayout(size_t nseg, size_t cap, int B, size_t N, int C) {
        size_t o = 0;
        counts = o; o += align256(4 * nseg);
        offsets = o; o += align256(4 * (nseg + 1));
        cursor = o; o += align256(4 * nseg);
        kept = o; o += align256(4 * nseg);
        keys = o; o += align256(8 * (cap ? cap : 1));
        mask_words = (N * (size_t)C + 31) / 32;       // one bit per score of an image (class-specific scan)
        mask = o; o += align256(4 * mask_words * (size_t)(B > 0 ? B : 0));
        total = o;
    }
};

}  // namespace effdet

using namespace effdet;

extern "C" int effdet_regress_boxes(const float *anchors, int per_image, const float *deltas,
                                    const float mean[4], const float stdv[4], int B, size_t N,
                                    float *out, void *stream) {
    return launch_boxes<true, false>(anchors, per_image, deltas, mean, stdv, B, N, 0, 0, out, stream);
}
extern "C" int effdet_clip_boxes(const float *boxes, int B, size_t N, float height, float width,
                                 float *out, void *stream) {
    return launch_boxes<false, true>(nullptr, 1, boxes, nullptr, nullptr, B, N, height, width, out,
                                     stream);
}
extern "C" int effdet_regress_clip_boxes(const float *anchors, int per_image, const float *deltas,
                                         const float mean[4], const float stdv[4], int B, size_t N,
                                         float height, float width, float *out, void *stream) {
    return launch_boxes<true, true>(anchors, per_image, deltas, mean, stdv, B, N, height, width, out,
                                    stream);
}

extern "C" size_t effdet_filter_detections_workspace_size(int B, size_t N, int C,
                                                          size_t cand_capacity, int max_det) {
    (void)max_det;
    if (B <= 0 || C <= 0) return 256;
    return FdLayout((size_t)B * C, cand_capacity, B, N, C).total;
}

extern "C" int effdet_filter_detections(const float *boxes, const float *classification, int B,
                                        size_t N, int C, float score_threshold, float iou_threshold,
                                        int max_det, int class_specific, int nms, void *workspace,
                                        size_t workspace_bytes, size_t cand_capacity,
                                        float *out_boxes, float *out_scores, int32_t *out_labels,
                                        int32_t *out_indices, int32_t *status, void *stream) {
    EFFDET_REQUIRE(B >= 0 && C >= 1 && max_det >= 1, "bad sizes");
    if (B == 0) return EFFDET_OK;
    EFFDET_REQUIRE(boxes && classification && workspace && out_boxes && out_scores && out_labels &&
                       status, "null pointer");
    EFFDET_REQUIRE(aligned16(boxes) && aligned16(out_boxes), "boxes must be 16-byte aligned");
    EFFDET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    EFFDET_REQUIRE((double)N * C < 4294967295.0, "N*C must fit in 32 bits");
    EFFDET_REQUIRE(cand_capacity < 0xffffffffull, "cand_capacity must fit in 32 bits");
    // reference: `if iou_threshold > 0` decides whether NMS runs (FilterDetections.py:11);
    // the layer forces the threshold to 0 when nms=False (:170-171)
    const int do_nms = (nms && iou_threshold > 0.f) ? 1 : 0;
    EFFDET_REQUIRE(!do_nms || max_det <= 2048, "max_detections > 2048 unsupported with NMS");
    const int S = class_specific ? C : 1;
    const size_t nseg = (size_t)B * S;
    FdLayout L(nseg, cand_capacity, B, N, C);
    if (workspace_bytes < L.total)
        return fail(EFFDET_E_CAPACITY, "effdet_filter_detections: workspace %s%lld < %lld bytes", "",
                    (long long)workspace_bytes, (long long)L.total);
    char *ws = static_cast<char *>(workspace);
    uint32_t *counts = reinterpret_cast<uint32_t *>(ws + L.counts);
    uint32_t *offsets = reinterpret_cast<uint32_t *>(ws + L.offsets);
    uint32_t *cursor = reinterpret_cast<uint32_t *>(ws + L.cursor);
    uint32_t *kept = reinterpret_cast<uint32_t *>(ws + L.kept);
    u64 *keys = reinterpret_cast<u64 *>(ws + L.keys);
    cudaStream_t st = as_stream(stream);
    const float4 *b4 = reinterpret_cast<const float4 *>(boxes);

    EFFDET_CUDA(cudaMemsetAsync(counts, 0, 4 * nseg, st));
    if (N > 0) {
        if (class_specific) {
            const uint32_t NC = (uint32_t)(N * C);
            dim3 grid(cdiv(NC, kScanElems), B);
            size_t sm = 2 * (size_t)C * sizeof(uint32_t);
            EFFDET_REQUIRE(sm <= 48 * 1024, "too many classes");
            uint32_t *mask = reinterpret_cast<uint32_t *>(ws + L.mask);
            EFFDET_CUDA(cudaMemsetAsync(mask, 0, 4 * L.mask_words * (size_t)B, st));
            scan_scores_kernel<<<grid, kScanThreads, sm, st>>>(classification, NC, C, score_threshold, counts, mask,
                                                               (uint32_t)L.mask_words);
            EFFDET_LAUNCHED();
            offsets_kernel<<<1, 1024, 0, st>>>(counts, (uint32_t)nseg, offsets, cursor,
                                               (u64)cand_capacity, status);
            EFFDET_LAUNCHED();
            fill_keys_kernel<<<dim3(cdiv(L.mask_words, kFillWords), B), kScanThreads, sm, st>>>(
                classification, NC, C, offsets, cursor, keys, status, mask, (uint32_t)L.mask_words);
            EFFDET_LAUNCHED();
        } else {
            dim3 grid(cdiv(N, kScanThreads), B);
            scan_rows_kernel<false><<<grid, kScanThreads, 0, st>>>(
                classification, (uint32_t)N, C, score_threshold, counts, offsets, cursor, keys, status);
            EFFDET_LAUNCHED();
            offsets_kernel<<<1, 1024, 0, st>>>(counts, (uint32_t)nseg, offsets, cursor,
                                               (u64)cand_capacity, status);
            EFFDET_LAUNCHED();
            scan_rows_kernel<true><<<grid, kScanThreads, 0, st>>>(
                classification, (uint32_t)N, C, score_threshold, counts, offsets, cursor, keys, status);
            EFFDET_LAUNCHED();
        }
    } else {
        offsets_kernel<<<1, 1024, 0, st>>>(counts, (uint32_t)nseg, offsets, cursor,
                                           (u64)cand_capacity, status);
        EFFDET_LAUNCHED();
    }
    const size_t sel_bytes = do_nms ? (size_t)max_det * sizeof(float4) : 0;
    if (sel_bytes > 24 * 1024) {
        // sort_nms_small_kernel holds 19 KiB of static shared memory: with the selected-box list of more than ~1850
        // detections (max_detections is validated up to 2048) static + dynamic passes the 48 KiB a launch gets
        // without opting in
        static bool sel_attr = false;
        if (!sel_attr) {
            EFFDET_CUDA(cudaFuncSetAttribute(sort_nms_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             2048 * (int)sizeof(float4)));
            EFFDET_CUDA(cudaFuncSetAttribute(sort_nms_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             2048 * (int)sizeof(float4)));
            sel_attr = true;
        }
    }
    sort_nms_small_kernel<<<(unsigned)nseg, kSmallThreads, sel_bytes, st>>>(
        b4, N, keys, offsets, counts, kept, S, iou_threshold, (uint32_t)max_det, do_nms, status);
    EFFDET_LAUNCHED();
    sort_nms_large_kernel<<<(unsigned)nseg, kLargeThreads, sel_bytes, st>>>(
        b4, N, keys, offsets, counts, kept, S, iou_threshold, (uint32_t)max_det, do_nms, status);
    EFFDET_LAUNCHED();
    const size_t merge_smem = ((size_t)S * max_det + max_det) * 4;
    if (merge_smem <= 200 * 1024 && max_det <= 65535 && S <= kMergeMaxLists) {
#define MERGE_LAUNCH(NQ)                                                                                          \
        {                                                                                                         \
            static bool attr = false;                                                                             \
            if (!attr) {                                                                                          \
                EFFDET_CUDA(cudaFuncSetAttribute(merge_topk_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 200 * 1024));                                                    \
                attr = true;                                                                                      \
            }                                                                                                     \
            merge_topk_kernel<NQ><<<B, kMergeThreads, merge_smem, st>>>(                                          \
                b4, classification, B, N, C, S, keys, offsets, kept, max_det,                                     \
                reinterpret_cast<float4 *>(out_boxes), out_scores, out_labels, out_indices, status);              \
        }
        if (S <= 32) MERGE_LAUNCH(1)
        else if (S <= 96) MERGE_LAUNCH(3)
        else MERGE_LAUNCH(8)
#undef MERGE_LAUNCH
    } else {
        merge_topk_warp_kernel<<<cdiv(B, kMergeWarps), kMergeWarps * 32, (size_t)kMergeWarps * S * 4, st>>>(
            b4, classification, B, N, C, S, keys, offsets, kept, max_det,
            reinterpret_cast<float4 *>(out_boxes), out_scores, out_labels, out_indices, status);
    }
    EFFDET_LAUNCHED();
    return EFFDET_OK;
}
