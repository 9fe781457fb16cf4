// BN254 G1 variable-base MSM on one B200: the kernels and the launch sequence.
//
// Replaces /root/reference/plonkish_backend/src/util/arithmetic/msm.rs:84-181
// (variable_base_msm + variable_base_msm_serial).  The reference walks unsigned
// c-bit windows high->low per CPU thread and does a random read-modify-write on
// a bucket array per (point, window) (msm.rs:168-173).  Here the same sum is
// reorganised for HBM and the integer pipe, in two bucket layouts (MsmPlan::mode):
//
//   mode 0  plain bases: one bucket set per window (c <= 16), window weights 2^(c*w)
//           applied at the end (the doubling chain of msm.rs:162-164).
//   mode 1  resident bases expanded once into T[w][i] = 2^(c*w) * P_i: every window feeds
//           ONE bucket set (c up to 22), no doubling chain, fewer windows per point.
//
//   K1 decompose   Fr Montgomery -> canonical (to_repr, msm.rs:153), negate if
//                  > (r-1)/2, signed base-2^c digits; histogram of the high bucket bits
//                  in shared memory.                                  [HBM bound]
//   K2 sort        two-level counting sort of (bucket, point) pairs on the high, then the
//                  low bucket bits.  mode 1: ranks in shared memory, tiles staged sorted,
//                  one global atomic per (block, digit) run, coalesced run writes, bins cut
//                  into slices shared by several blocks.  Output: one dense array ordered by
//                  bucket and bucket_start[].                          [HBM bound]
//   K3 accumulate  every thread owns a fixed-length run of the sorted array (skew
//                  proof), sums its points with XYZZ mixed additions, stores the
//                  buckets that lie wholly inside the run and leaves its two open ends as
//                  items; warp-level segmented reductions by shuffles fold the items,
//                  16-128x fewer per level.                           [IMAD bound]
//   K4 bucket reduce   sum_k k*B_k by chunked running sums (msm.rs:175-179 per chunk).
//   K5 window combine  2^(c*w) weights (mode 0), final sum, XYZZ -> affine.
//
// This header is also compiled by g++ against tests/emul/cuda_emul.h so the
// kernels' logic runs in the CPU test suite; the product only runs the nvcc build.
#pragma once
#include "g1.cuh"

#ifndef PLONKISH_EMUL
#include <cuda_runtime.h>
#ifndef PK_COUNT_LAUNCH
#define PK_COUNT_LAUNCH() ((void)0)
#endif
#define PK_LAUNCH(kern, grid, block, smem, stream, ...) \
    do { PK_COUNT_LAUNCH(); kern<<<grid, block, smem, stream>>>(__VA_ARGS__); } while (0)
#define PK_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; type *name = reinterpret_cast<type *>(name##_raw)
typedef cudaStream_t pk_stream_t;
#define PK_MARK(marks, i, stream) do { if (marks) cudaEventRecord((cudaEvent_t)(marks)->ev[i], stream); } while (0)
#define PK_MEMSET0(ptr, bytes, stream) cudaMemsetAsync(ptr, 0, bytes, stream)
#define PK_SET_SMEM(kern, bytes) do { if ((bytes) > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); } while (0)
#else
#define PK_MEMSET0(ptr, bytes, stream) memset(ptr, 0, bytes)
#define PK_SET_SMEM(kern, bytes) ((void)0)
#define PK_MARK(marks, i, stream) ((void)0)
#define PK_LAUNCH(kern, grid, block, smem, stream, ...) emul::launch(grid, block, smem, [&] { kern(__VA_ARGS__); })
#define PK_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emul::g_dyn_smem)
typedef int pk_stream_t;
#endif

namespace pk {

typedef uint16_t u16;

static const u32 PK_INVALID_KEY = 0xffffffffu;
static const u32 PK_ZERO_DIGIT = 0xffffu;

// ------------------------------------------------------------------ the plan
struct MsmPlan {
    u32 n;          // points in this launch sequence (<= 2^26)
    u32 n_pad;      // n rounded up to 8 (digit rows are read as uint4)
    u32 c;          // window bits, 8..16
    u32 W;          // windows = ceil(254 / c) (scalars are folded to <= (r-1)/2 first)
    u32 B;          // buckets per window = 2^(c-1)
    u32 hi_bits, lo_bits, idx_bits;
    u32 HI;         // 2^hi_bits bins per window
    u32 nbins;      // W * HI
    u32 nbuckets;   // W * B
    u32 tile;       // points per K1 block
    u32 ntiles;
    u32 L;          // run length per K3 thread
    u32 nthreads1;  // K3 threads (multiple of blk_acc)
    u32 blk_acc;    // K3 block size: 128, or 32 when one warp covers everything
    u32 rb;         // buckets per K4 thread
    u32 red_threads;  // K4 threads per window
    u32 red_blocks;   // K4 blocks per window
    u32 blk;          // threads per block of the streaming kernels (256; tests shrink it)
    u32 serial_items; // item levels above this many items let every lane fold 8 items serially first
    u32 blk_stage;    // threads per block of the staged partition kernels (512, 1024 from 2^24 points; tests shrink it)
    u32 mode;         // 0: one bucket set per window (any bases); 1: one bucket set, bases are a table of
                      //    precomputed window multiples T[w][i] = 2^(c*w) * P_i (resident bases only)
    u32 ngroups;      // bucket sets: W in mode 0, 1 in mode 1
    u32 stride;       // mode 1: points per table row
    u32 chunk;        // which of the MSM's bucket arrays this launch sequence fills (host path pipelining)
    u32 nchunks;      // bucket arrays the reduce phase adds up (1 unless the host path cut the points into chunks)
    u32 fuse_l1;      // mode 1, c >= 16: level 1 of the sort recomputes the digits from the scalars (no digit array)
    u32 acc_tiers;    // K3 run lengths: 0 = every thread takes L entries; J > 0 = up to J tiers of decreasing run length (pk_acc_run)
    u32 acc_resident; // K3 threads the device holds at once (sm_count * 512)
};

inline u32 pk_ceil_log2(u32 v) {
    u32 b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

inline u32 pk_default_window_bits(u32 n) {
    u32 lg = pk_ceil_log2(n < 2 ? 2 : n);
    if (lg >= 22) return 16;
    if (lg >= 20) return 15;
    if (lg >= 18) return 14;
    if (lg >= 16) return 13;
    if (lg >= 14) return 12;
    if (lg >= 12) return 10;
    return 8;
}

// Bucket reduce geometry: thread j owns buckets [j*rb, (j+1)*rb) of a group.  Every thread pays a
// fixed ~400 field products (its small scalar multiplication and the warp sum) on top of 28 per
// bucket, so rb grows with the bucket count until the groups together fill the GPU with about two
// warps per scheduler; rb need not divide B (the last thread's range is cut).
#define PK_RED_BLOCK 128
inline void pk_plan_reduce(MsmPlan &p, u32 sm_count) {
    const unsigned long long target = (unsigned long long)sm_count * 4ull * 32ull * 2ull;  // threads over all groups
    const unsigned long long groups = p.ngroups ? p.ngroups : 1;
    const unsigned long long total = (unsigned long long)p.B * groups;
    unsigned long long rb = (total + target - 1) / target;
    // small bucket counts are latency bound: a shorter serial chain per thread wins (measured: 2^16 buckets 0.34 -> 0.29 ms
    // with 4 per thread, 2^14 0.32 -> 0.25 ms and 2^10 0.27 -> 0.20 ms with 2; from 2^17 buckets on 8 is best)
    const unsigned long long floor_rb = total >= (1ull << 17) ? 8 : total >= (1ull << 15) ? 4 : 2;
    if (rb < floor_rb) rb = floor_rb;
    if (rb > p.B) rb = p.B;
    p.rb = (u32)rb;
    p.red_threads = (p.B + p.rb - 1) / p.rb;
    p.red_blocks = (p.red_threads + PK_RED_BLOCK - 1) / PK_RED_BLOCK;
}

// ---------------------------------------------------------------- K3 run lengths
// Thread t of k_accumulate sums one run of consecutive sorted entries and leaves two items for the segmented
// reduction.  Every run costs about 3.5 mixed additions on top of its entries (the items), and a launch of equal runs
// ends with a ragged last wave: a slot that frees up late finds no work, about half a run's duration per slot
// (measured on B200, tools/acc_run_length.py: the accumulate of 2^21 points takes 4.28 ms with 3 waves of 119-entry
// runs, 3.65 ms with 24 waves of 16 — but then the items cost 0.64 ms instead of 0.14).  So the runs shrink as the launch
// proceeds (guided self-scheduling; blocks are dispatched in index order): tier j takes half of what is left in runs
// of rem / (2 R) entries — one wave of the R resident threads — until the runs reach 16 entries or the last of J tiers
// takes the rest.  About (J + 1) R threads in all, a tail of 2^-(J+1) of the work.  The tiers depend on the actual
// entry count (zero digits are dropped, skewed scalars drop most of them), so every thread derives its run from
// `total` on the device; the host only sizes the launch (pk_acc_threads) for the largest count.
#define PK_ACC_LMIN 16u
#define PK_ACC_LMAX 1024u
#ifndef PLONKISH_EMUL
#define PK_HOST_DEVICE __host__ __device__ __forceinline__
#else
#define PK_HOST_DEVICE inline
#endif
PK_HOST_DEVICE u32 pk_acc_tier_len(u32 rem, u32 resident) {
    const u32 r2 = 2u * resident;
    u32 L = (rem + r2 - 1u) / r2;
    L = (L + 7u) & ~7u;  // multiples of 8 keep the lanes' reads of `sorted` on sector boundaries
    if (L < PK_ACC_LMIN) L = PK_ACC_LMIN;
    if (L > PK_ACC_LMAX) L = PK_ACC_LMAX;
    return L;
}
// The run [s, e) of thread t among nthreads for `total` entries; false when the thread has none.
PK_HOST_DEVICE bool pk_acc_run(u32 t, u32 total, u32 tiers, u32 resident, u32 nthreads, u32 uniform_L, u32 &s, u32 &e) {
    if (tiers == 0) {
        const unsigned long long s64 = (unsigned long long)t * uniform_L;
        if (s64 >= total) return false;
        s = (u32)s64;
        e = (total - s < uniform_L) ? total : s + uniform_L;
        return true;
    }
    u32 start = 0, tid = t, avail = nthreads;
    for (u32 j = 0; j < tiers; ++j) {
        const u32 rem = total - start;
        if (rem == 0 || avail == 0) return false;
        u32 L = pk_acc_tier_len(rem, resident);
        const u32 nb = (rem / 2u) / (128u * L);  // whole 128-thread blocks, so no idle thread sits between two runs
        if (j + 1 == tiers || L == PK_ACC_LMIN || nb == 0 || nb * 128u >= avail) {
            // last tier: everything that is left, in runs long enough for the threads that are left
            u32 need = (rem + avail - 1u) / avail;
            need = (need + 7u) & ~7u;
            if (need > L) L = need;
            const unsigned long long off = (unsigned long long)tid * L;
            if (off >= rem) return false;
            s = start + (u32)off;
            e = (rem - (u32)off < L) ? total : s + L;
            return true;
        }
        const u32 nt = nb * 128u;
        if (tid < nt) {
            s = start + tid * L;
            e = s + L;
            return true;
        }
        tid -= nt;
        avail -= nt;
        start += nt * L;
    }
    return false;
}
// Threads to launch: what the tiers take for the largest entry count (the same walk as pk_acc_run) plus a block per
// tier of slack.  A smaller count whose rounding asks for a few threads more is still covered: the final tier
// stretches its runs over the threads that are left.  Idle threads cost two (empty) items each, so no generous bound.
inline u32 pk_acc_threads(unsigned long long entries, u32 tiers, u32 resident) {
    unsigned long long nt = 0, start = 0;
    for (u32 j = 0; j < tiers; ++j) {
        const unsigned long long rem = entries - start;
        if (rem == 0) break;
        const u32 L = pk_acc_tier_len((u32)(rem > 0xffffffffull ? 0xffffffffull : rem), resident);
        const unsigned long long nb = (rem / 2ull) / (128ull * L);
        if (j + 1 == tiers || L == PK_ACC_LMIN || nb == 0) {
            nt += (rem + L - 1) / L;
            break;
        }
        nt += nb * 128ull;
        start += nb * 128ull * L;
    }
    nt = ((nt + 127ull) & ~127ull) + 128ull * (tiers + 1);
    return (u32)nt;
}

// Launch geometry of K3.  L > 0: equal runs of L entries (tuning / tests); L = 0: the tiers above.
#define PK_ACC_TIERS 7u
inline void pk_set_run_length(MsmPlan &p, u32 L) {
    const unsigned long long emax = (unsigned long long)p.n * p.W;
    if (L == 0 && emax > 32ull * PK_ACC_LMIN) {
        p.acc_tiers = PK_ACC_TIERS;
        p.L = pk_acc_tier_len((u32)emax, p.acc_resident);  // first tier's run for the largest entry count (reported by msm_plan)
        p.nthreads1 = pk_acc_threads(emax, p.acc_tiers, p.acc_resident);
        p.blk_acc = 128u;
        return;
    }
    p.acc_tiers = 0;
    p.L = L ? L : PK_ACC_LMIN;
    const unsigned long long t1 = (emax + p.L - 1) / p.L;
    // one warp -> a 32-thread terminal launch; otherwise whole 128-thread blocks
    p.nthreads1 = (t1 <= 32) ? 32u : (u32)((t1 + 127ull) & ~127ull);
    p.blk_acc = (t1 <= 32) ? 32u : 128u;
}

inline MsmPlan pk_make_plan(u32 n, u32 c_override, u32 sm_count) {
    MsmPlan p;
    p.n = n;
    p.n_pad = (n + 7u) & ~7u;
    p.c = c_override ? c_override : pk_default_window_bits(n);
    if (p.c < 8) p.c = 8;
    if (p.c > 16) p.c = 16;
    p.W = (254 + p.c - 1) / p.c;
    if (p.c * p.W < 255) p.W += 1;  // keep one spare bit so the top digit never carries out
    p.B = 1u << (p.c - 1);
    p.idx_bits = pk_ceil_log2(n < 2 ? 2 : n);
    u32 hi = (p.c - 1 < 8) ? p.c - 1 : 8;
    u32 room = 31 - p.idx_bits;  // bits left for the low bucket bits in a scatter entry
    if (p.c - 1 > hi + room) hi = p.c - 1 - room;
    p.hi_bits = hi;
    p.lo_bits = p.c - 1 - hi;
    p.HI = 1u << hi;
    p.nbins = p.W * p.HI;
    p.nbuckets = p.W * p.B;
    u32 tile = 1024;
    while (tile < 65536 && (n + tile - 1) / tile > 1024) tile <<= 1;
    p.tile = tile;
    p.ntiles = (n + tile - 1) / tile;
    p.acc_resident = sm_count * 512u;
    pk_set_run_length(p, 0);
    p.ngroups = p.W;
    pk_plan_reduce(p, sm_count);
    p.blk = 256;
    p.serial_items = 8192;
    p.blk_stage = 512;
    p.mode = 0;
    p.ngroups = p.W;
    p.stride = 0;
    p.chunk = 0;
    p.nchunks = 1;
    p.fuse_l1 = n <= (1u << 21) ? 1u : 0u;  // see plan_for in api.cu: the fused level 1 wins on small launches only
    return p;
}

// Window width used when a resident base slice of n_registered points is expanded
// into its table of window multiples.  One bucket set serves all windows, so the reduce
// costs ~30-40 field products per bucket against 10 per (point, window) in the accumulate:
// measured on B200 (tools/table_c_sweep.py, 2^16..2^24) the total is flat or best at
// c = log2(n) - 1 (n/4 buckets, ~50 entries per bucket), capped at 22 by the sort's digit widths.
inline u32 pk_table_window_bits(u32 n_registered) {
    int c = (int)pk_ceil_log2(n_registered < 2 ? 2 : n_registered) - 1;
    if (c < 8) c = 8;
    if (c > 22) c = 22;
    return (u32)c;
}

inline u32 pk_windows_for(u32 c) {
    u32 w = (254 + c - 1) / c;
    if (c * w < 255) w += 1;
    return w;
}

// Mode 1 plan: n points (a prefix of a table with `stride` points per row), c fixed by the table.
inline MsmPlan pk_make_plan_b(u32 n, u32 c, u32 stride, u32 sm_count) {
    MsmPlan p = pk_make_plan(n, 16, sm_count);
    p.mode = 1;
    p.c = c;
    p.W = pk_windows_for(c);
    p.B = 1u << (c - 1);
    p.idx_bits = 0;
    p.hi_bits = (c - 1 < 10) ? c - 1 : 10;
    p.lo_bits = c - 1 - p.hi_bits;
    p.HI = 1u << p.hi_bits;
    p.nbins = p.HI;
    p.nbuckets = p.B;
    p.ngroups = 1;
    p.stride = stride;
    // 1 024-thread stages (16 K entries) double the run length of both sort levels; they pay once there are
    // enough tiles to fill the GPU with one such block per SM (measured at 2^24: scatter 1.22 -> 1.03 ms,
    // level 2 1.35 -> 1.27 ms; smaller MSMs have tiles shorter than such a stage: 2^22 0.41 -> 0.55 ms)
    p.blk_stage = (p.tile >= 16384) ? 1024 : 512;  // a tile row holds at least one 16 K-entry stage: n > 2^23
    pk_set_run_length(p, 0);
    pk_plan_reduce(p, sm_count);
    return p;
}

// Workspace layout (one arena, 256-byte aligned pieces).
struct MsmWorkspace {
    u16 *digits;        // [W][n_pad]           (mode 1: u32 digits in the same storage)
    u32 *tile_hist;     // [ntiles][nbins]
    u32 *bin_total;     // [nbins]
    u32 *bin_start;     // [nbins + 1]
    u32 *slice_prefix;  // [nbins + 1] mode 1: first level-2 slice of every bin
    u32 *l1;            // [n * W] entries after the bin scatter (mode 1: u64 entries)
    u32 *sorted;        // [n * W] (sign << 31 | point index), ordered by (window, bucket)
    u32 *bucket_start;  // [nbuckets + 1]
    u32 *bucket_cur;    // [nbuckets + 1] mode 1: per-bucket counts, then cursors
    u32 *scan_tmp;      // [2 * 2050] block totals / offsets of the multi-block scan
    xyzz *bucket_sum;   // [nbuckets]
    u32 *item_keys[2];  // ping-pong item lists for the segmented reduction levels
    xyzz *item_pts[2];
    xyzz *block_out;    // [W][red_blocks]
    xyzz *win_out;      // [W] weighted window sums
    xyzz *result;       // [1] projective result of this launch sequence
};

inline size_t pk_align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Pieces whose size depends only on (mode, c) come first, so that every chunk of one MSM
// (same table, different point counts) finds bucket_sum etc. at the same offsets.
inline size_t pk_workspace_bytes(const MsmPlan &p) {
    size_t e = (size_t)p.n * p.W;
    size_t items = (size_t)p.nthreads1 * 2;
    size_t s = 0;
    s += pk_align256(sizeof(xyzz) * (size_t)p.nbuckets * p.nchunks);
    s += 2 * pk_align256(sizeof(u32) * (p.nbuckets + 1));
    s += pk_align256(sizeof(u32) * 2 * 2050);
    s += pk_align256(sizeof(xyzz) * (size_t)p.ngroups * p.red_blocks);
    s += pk_align256(sizeof(xyzz) * 32);
    s += pk_align256(sizeof(xyzz));
    s += pk_align256(sizeof(u32) * p.nbins);
    s += 2 * pk_align256(sizeof(u32) * (p.nbins + 1));
    s += pk_align256((p.mode ? sizeof(u32) : sizeof(u16)) * (size_t)p.W * p.n_pad);
    s += pk_align256(sizeof(u32) * (size_t)p.ntiles * p.nbins);
    s += pk_align256((p.mode ? sizeof(unsigned long long) : sizeof(u32)) * e);
    s += pk_align256(sizeof(u32) * e);
    s += 2 * pk_align256(sizeof(u32) * items);
    s += 2 * pk_align256(sizeof(xyzz) * items);
    return s;
}

inline MsmWorkspace pk_carve_workspace(const MsmPlan &p, void *arena) {
    MsmWorkspace w;
    unsigned char *q = (unsigned char *)arena;
    size_t e = (size_t)p.n * p.W;
    size_t items = (size_t)p.nthreads1 * 2;
    w.bucket_sum = (xyzz *)q; q += pk_align256(sizeof(xyzz) * (size_t)p.nbuckets * p.nchunks);
    w.bucket_start = (u32 *)q; q += pk_align256(sizeof(u32) * (p.nbuckets + 1));
    w.bucket_cur = (u32 *)q; q += pk_align256(sizeof(u32) * (p.nbuckets + 1));
    w.scan_tmp = (u32 *)q; q += pk_align256(sizeof(u32) * 2 * 2050);
    w.block_out = (xyzz *)q; q += pk_align256(sizeof(xyzz) * (size_t)p.ngroups * p.red_blocks);
    w.win_out = (xyzz *)q; q += pk_align256(sizeof(xyzz) * 32);
    w.result = (xyzz *)q; q += pk_align256(sizeof(xyzz));
    w.bin_total = (u32 *)q; q += pk_align256(sizeof(u32) * p.nbins);
    w.bin_start = (u32 *)q; q += pk_align256(sizeof(u32) * (p.nbins + 1));
    w.slice_prefix = (u32 *)q; q += pk_align256(sizeof(u32) * (p.nbins + 1));
    w.digits = (u16 *)q; q += pk_align256((p.mode ? sizeof(u32) : sizeof(u16)) * (size_t)p.W * p.n_pad);
    w.tile_hist = (u32 *)q; q += pk_align256(sizeof(u32) * (size_t)p.ntiles * p.nbins);
    w.l1 = (u32 *)q; q += pk_align256((p.mode ? sizeof(unsigned long long) : sizeof(u32)) * e);
    w.sorted = (u32 *)q; q += pk_align256(sizeof(u32) * e);
    for (int k = 0; k < 2; ++k) { w.item_keys[k] = (u32 *)q; q += pk_align256(sizeof(u32) * items); }
    for (int k = 0; k < 2; ++k) { w.item_pts[k] = (xyzz *)q; q += pk_align256(sizeof(xyzz) * items); }
    return w;
}

// ------------------------------------------------------- 128-bit load/store
PK_HD fe load_fe(const uint4 *p) {
    uint4 a = __ldg(p), b = __ldg(p + 1);
    fe r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
    return r;
}
PK_HD void store_fe(uint4 *p, const fe &v) {
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
PK_HD xyzz load_xyzz(const xyzz *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    xyzz r;
    r.x = load_fe(q); r.y = load_fe(q + 2); r.zz = load_fe(q + 4); r.zzz = load_fe(q + 6);
    return r;
}
PK_HD void store_xyzz(xyzz *p, const xyzz &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    store_fe(q, v.x); store_fe(q + 2, v.y); store_fe(q + 4, v.zz); store_fe(q + 6, v.zzz);
}

// ================================================================= K1 decompose
// (r - 1) / 2, little-endian 32-bit limbs.
PK_HD bool fr_above_half(const fe &v) {
    const u32 h[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (v.l[i] != h[i]) return v.l[i] > h[i];
    }
    return false;
}

template <int C>
__global__ void __launch_bounds__(256) k_decompose(const uint4 *__restrict__ scalars, MsmPlan p, u16 *__restrict__ digits,
                                                   u32 *__restrict__ tile_hist) {
    constexpr int W = (254 + C - 1) / C + ((C * ((254 + C - 1) / C) < 255) ? 1 : 0);
    constexpr u32 B = 1u << (C - 1);
    PK_DYN_SMEM(u32, hist);  // [W][HI]
    const u32 tile = blockIdx.x;
    for (u32 k = threadIdx.x; k < p.nbins; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const u32 beg = tile * p.tile;
    const u32 end = (beg + p.tile < p.n) ? beg + p.tile : p.n;
    for (u32 i = beg + threadIdx.x; i < end; i += blockDim.x) {
        fe v = fr_to_canonical(load_fe(scalars + 2 * (size_t)i));
        const bool neg = fr_above_half(v);
        if (neg) {
            u32 m[8];
            FrMod::limbs(m);
            fe t;
            sub8(t.l, m, v.l);
            v = t;
        }
        u32 carry = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int off = w * C;
            const int word = off >> 5, sh = off & 31;
            u32 raw = 0;
            if (word < 8) {
                raw = v.l[word] >> sh;
                if (sh + C > 32 && word + 1 < 8) raw |= v.l[word + 1] << (32 - sh);
            }
            raw = (raw & ((1u << C) - 1u)) + carry;
            // Non-negated scalars borrow when raw > B, negated ones when raw >= B, so
            // the encoded (sign, |d| - 1) pair never collides with PK_ZERO_DIGIT.
            const bool borrow = neg ? (raw >= B) : (raw > B);
            const u32 mag = borrow ? (1u << C) - raw : raw;
            carry = borrow ? 1u : 0u;
            const u32 sign = (borrow ? 1u : 0u) ^ (neg ? 1u : 0u);
            u32 enc = PK_ZERO_DIGIT;
            if (mag != 0) {
                enc = (sign << 15) | (mag - 1u);
                atomicAdd(&hist[w * p.HI + ((mag - 1u) >> p.lo_bits)], 1u);
            }
            digits[(size_t)w * p.n_pad + i] = (u16)enc;
        }
    }
    __syncthreads();
    u32 *out = tile_hist + (size_t)tile * p.nbins;
    for (u32 k = threadIdx.x; k < p.nbins; k += blockDim.x) out[k] = hist[k];
}

// ===================================================================== K2 scans
// Column scan: for each bin, exclusive prefix over tiles (in place) and the total.
__global__ void __launch_bounds__(256) k_scan_tiles(u32 *__restrict__ tile_hist, u32 ntiles, u32 nbins, u32 *__restrict__ bin_total) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    u32 run = 0;
    for (u32 t = 0; t < ntiles; ++t) {
        const u32 v = tile_hist[(size_t)t * nbins + b];
        tile_hist[(size_t)t * nbins + b] = run;
        run += v;
    }
    bin_total[b] = run;
}

// Exclusive scan of bin_total[nbins] into bin_start[nbins + 1]; one block of 1024.
__global__ void __launch_bounds__(1024) k_scan_bins(const u32 *__restrict__ bin_total, u32 nbins, u32 *__restrict__ bin_start) {
    __shared__ u32 part[1024];
    const u32 per = (nbins + blockDim.x - 1) / blockDim.x;
    const u32 beg = threadIdx.x * per;
    u32 sum = 0;
    for (u32 k = 0; k < per; ++k) {
        if (beg + k < nbins) sum += bin_total[beg + k];
    }
    part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the per-thread sums.
    for (u32 d = 1; d < blockDim.x; d <<= 1) {
        u32 v = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    u32 run = part[threadIdx.x] - sum;
    for (u32 k = 0; k < per; ++k) {
        if (beg + k < nbins) {
            bin_start[beg + k] = run;
            run += bin_total[beg + k];
        }
    }
    if (threadIdx.x == blockDim.x - 1) bin_start[nbins] = part[threadIdx.x];
}

// =========================================================== K2 bin scatter (L1)
// grid (ntiles, W).  Entry layout: lo_bits | sign | point index (idx_bits).
__global__ void __launch_bounds__(256) k_scatter_bins(const u16 *__restrict__ digits, MsmPlan p, const u32 *__restrict__ tile_hist,
                                                      const u32 *__restrict__ bin_start, u32 *__restrict__ l1) {
    __shared__ u32 cursor[1024];
    const u32 tile = blockIdx.x, w = blockIdx.y;
    for (u32 h = threadIdx.x; h < p.HI; h += blockDim.x)
        cursor[h] = bin_start[w * p.HI + h] + tile_hist[(size_t)tile * p.nbins + w * p.HI + h];
    __syncthreads();
    const u32 beg = tile * p.tile;
    const u32 end = (beg + p.tile < p.n) ? beg + p.tile : p.n;
    const u16 *row = digits + (size_t)w * p.n_pad;
    const u32 lo_mask = (1u << p.lo_bits) - 1u;
    // tile is a multiple of 8 and rows are padded to 8: 128-bit loads of 8 digits, four
    // loads in flight per thread, then all their shared-memory cursor updates, then the stores.
    constexpr int U = 4;
    for (u32 i0 = beg + 8 * threadIdx.x; i0 < end; i0 += 8 * blockDim.x * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 i = i0 + (u32)u * 8 * blockDim.x;
            v[u] = (i < end) ? __ldg(reinterpret_cast<const uint4 *>(row + i)) : make_uint4(~0u, ~0u, ~0u, ~0u);
        }
        u32 pos[U * 8], ent[U * 8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 i = i0 + (u32)u * 8 * blockDim.x;
            const u32 words[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const u32 d = (words[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                const bool valid = (i + k < end) && d != PK_ZERO_DIGIT;
                const u32 b = d & 0x7fffu;
                ent[u * 8 + k] = ((b & lo_mask) << (p.idx_bits + 1)) | ((d >> 15) << p.idx_bits) | (i + k);
                pos[u * 8 + k] = valid ? atomicAdd(&cursor[b >> p.lo_bits], 1u) : 0xffffffffu;
            }
        }
#pragma unroll
        for (int j = 0; j < U * 8; ++j) {
            if (pos[j] != 0xffffffffu) l1[pos[j]] = ent[j];
        }
    }
}

// ============================================================= K2 bin sort (L2)
// One block per (window, high-bits) bin: counting sort on the low bucket bits.
// Writes bucket_start[] for the bin's buckets and the final (sign | index) entries.
__global__ void __launch_bounds__(256) k_sort_bins(const u32 *__restrict__ l1, MsmPlan p, const u32 *__restrict__ bin_start,
                                                   u32 *__restrict__ sorted, u32 *__restrict__ bucket_start) {
    __shared__ u32 cnt[128];
    const u32 bin = blockIdx.x;
    const u32 beg = bin_start[bin], end = bin_start[bin + 1];
    const u32 nlo = 1u << p.lo_bits;
    for (u32 k = threadIdx.x; k < nlo; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    const u32 sh = p.idx_bits + 1;
    constexpr int U = 8;  // independent loads in flight per thread
    for (u32 q0 = beg + threadIdx.x; q0 < end; q0 += blockDim.x * U) {
        u32 e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 q = q0 + (u32)u * blockDim.x;
            e[u] = (q < end) ? l1[q] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (q0 + (u32)u * blockDim.x < end) atomicAdd(&cnt[e[u] >> sh], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = beg;
        for (u32 k = 0; k < nlo; ++k) {
            const u32 v = cnt[k];
            cnt[k] = run;
            bucket_start[(size_t)bin * nlo + k] = run;
            run += v;
        }
        if (bin == p.nbins - 1) bucket_start[p.nbuckets] = run;
    }
    __syncthreads();
    const u32 idx_mask = (1u << p.idx_bits) - 1u;
    for (u32 q0 = beg + threadIdx.x; q0 < end; q0 += blockDim.x * U) {
        u32 e[U], pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 q = q0 + (u32)u * blockDim.x;
            e[u] = (q < end) ? l1[q] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pos[u] = (q0 + (u32)u * blockDim.x < end) ? atomicAdd(&cnt[e[u] >> sh], 1u) : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (pos[u] != 0xffffffffu) sorted[pos[u]] = (((e[u] >> p.idx_bits) & 1u) << 31) | (e[u] & idx_mask);
        }
    }
}

// ====================================================== mode 1: table of multiples
// With resident bases the windows share ONE bucket set: digit d of window w selects
// bucket |d| and the precomputed point T[w][i] = 2^(c*w) * P_i, so there is no
// per-window reduction and no 2^(c*w) doubling chain at the end (msm.rs:162-164
// disappears), and c can grow to 20-22 (fewer windows => fewer additions per point).
// STORE = false only counts (the fused level-1 scatter below recomputes the digits from the scalars instead of reading
// them back: 32 B per point read twice instead of 4 W bytes written and read).
template <int C, bool STORE = true>
__global__ void __launch_bounds__(256) k_decompose_b(const uint4 *__restrict__ scalars, MsmPlan p, u32 *__restrict__ digits,
                                                     u32 *__restrict__ tile_hist) {
    constexpr int W = (254 + C - 1) / C + ((C * ((254 + C - 1) / C) < 255) ? 1 : 0);
    constexpr u32 B = 1u << (C - 1);
    __shared__ u32 hist[1024];
    const u32 tile = blockIdx.x;
    for (u32 k = threadIdx.x; k < p.HI; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const u32 beg = tile * p.tile;
    const u32 end = (beg + p.tile < p.n) ? beg + p.tile : p.n;
    for (u32 i = beg + threadIdx.x; i < end; i += blockDim.x) {
        fe v = fr_to_canonical(load_fe(scalars + 2 * (size_t)i));
        const bool neg = fr_above_half(v);
        if (neg) {
            u32 m[8];
            FrMod::limbs(m);
            fe t;
            sub8(t.l, m, v.l);
            v = t;
        }
        u32 carry = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int off = w * C;
            const int word = off >> 5, sh = off & 31;
            u32 raw = 0;
            if (word < 8) {
                raw = v.l[word] >> sh;
                if (sh + C > 32 && word + 1 < 8) raw |= v.l[word + 1] << (32 - sh);
            }
            raw = (raw & ((1u << C) - 1u)) + carry;
            const bool borrow = neg ? (raw >= B) : (raw > B);  // see k_decompose
            const u32 mag = borrow ? (1u << C) - raw : raw;
            carry = borrow ? 1u : 0u;
            const u32 sign = (borrow ? 1u : 0u) ^ (neg ? 1u : 0u);
            u32 enc = 0xffffffffu;
            if (mag != 0) {
                enc = (sign << 31) | (mag - 1u);
                atomicAdd(&hist[(mag - 1u) >> p.lo_bits], 1u);
            }
            if (STORE) digits[(size_t)w * p.n_pad + i] = enc;
        }
    }
    __syncthreads();
    // tile_hist here is the global per-bin count (zeroed before the launch).
    for (u32 k = threadIdx.x; k < p.HI; k += blockDim.x) {
        if (hist[k]) atomicAdd(&tile_hist[k], hist[k]);
    }
}

// ---------------------------------------------------------------- staged partition
// Block-wide exclusive scan of cnt[0..R) into start[0..R); returns the total.  scratch: 33 words.
PK_HD u32 block_exclusive_scan(const u32 *cnt, u32 *start, u32 R, u32 *scratch) {
    const u32 per = (R + blockDim.x - 1) / blockDim.x;
    const u32 first = threadIdx.x * per;
    u32 sum = 0;
    for (u32 k = 0; k < per; ++k) {
        if (first + k < R) sum += cnt[first + k];
    }
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    u32 incl = sum;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const u32 nw = (blockDim.x + 31) >> 5;
        u32 w = (lane < nw) ? scratch[lane] : 0;
        u32 wi = w;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        scratch[lane] = wi - w;           // exclusive warp offsets
        if (lane == 31) scratch[32] = wi; // grand total
    }
    __syncthreads();
    u32 run = scratch[wid] + incl - sum;
    for (u32 k = 0; k < per; ++k) {
        if (first + k < R) {
            start[first + k] = run;
            run += cnt[first + k];
        }
    }
    const u32 total = scratch[32];
    __syncthreads();
    return total;
}

// Shared-memory view used by the staged kernels (dynamic smem):
//   u32 lcnt[R] | lstart[R] | gdelta[R] | scratch[64] | (pad to 8 bytes) | u64 stage[S]
// A staged entry is one 64-bit word (digit << 48 | aux << 32 | value): one random store when it is staged and one load
// when it is written out, instead of two each (the random shared accesses are what the stage phases wait for,
// tools/stage_phase_probe.py).  gdelta[d] = global base of the block's run for digit d minus its start in the stage.
struct StageSmem {
    u32 *lcnt, *lstart, *gdelta, *scratch;
    unsigned long long *stage;
};
PK_HD u32 stage_words_before(u32 R) { return (3u * R + 64u + 1u) & ~1u; }
PK_HD StageSmem stage_carve(u32 *base, u32 R, u32 S) {
    StageSmem m;
    (void)S;
    m.lcnt = base; m.lstart = base + R; m.gdelta = base + 2 * R; m.scratch = base + 3 * R;
    m.stage = reinterpret_cast<unsigned long long *>(base + stage_words_before(R));
    return m;
}
inline size_t stage_smem_bytes(u32 R, u32 S) { return sizeof(u32) * ((size_t)((3u * R + 64u + 1u) & ~1u) + 2 * (size_t)S); }

// L2 prefetch of a line the block reads in its next stage (the stage's own loads are issued right before their first
// use, and with one block per SM nothing else hides their DRAM latency).
PK_HD void prefetch_l2(const void *ptr) {
#ifndef PLONKISH_EMUL
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

// Optional phase timing of the staged partition (experiment builds only: -DPK_STAGE_PROF; tools/stage_phase_probe.py).
#ifdef PK_STAGE_PROF
__device__ unsigned long long g_stage_prof[16];
#define PK_PROF_START() unsigned long long pk_prof_t = clock64()
#define PK_PROF_ADD(k) do { if (blockIdx.x == 7 && threadIdx.x == 0) { const unsigned long long t1 = clock64(); atomicAdd(&g_stage_prof[k], t1 - pk_prof_t); pk_prof_t = t1; } } while (0)
#else
#define PK_PROF_START() ((void)0)
#define PK_PROF_ADD(k) ((void)0)
#endif

// Partition the block's EPT*blockDim items by digit: rank in shared memory, stage the
// items sorted by digit, claim one contiguous global range per (block, digit) run with a
// single atomicAdd on gcursor[digit], then write the runs out with consecutive lanes on
// consecutive addresses.  Needs lcnt zeroed on entry; leaves it zeroed.
// LATE: the values are not needed to rank; they are loaded from late_src[late_q + e * blockDim] once the ranks are taken,
// so that they are in flight during the scan instead of holding EPT registers through the rank phase.
template <int EPT, bool AUX = true, bool LATE = false>
PK_HD void staged_partition(u32 R, const u32 (&dig)[EPT], const u32 (&val_in)[EPT], const u32 (&aux)[EPT], const bool (&ok)[EPT],
                            const StageSmem &m, u32 *gcursor, u32 *out_val, u16 *out_aux, int prof_base = 0,
                            const u32 *late_src = nullptr, u32 late_q = 0) {
    PK_PROF_START();
    (void)prof_base;
    u32 rank[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) rank[e] = ok[e] ? atomicAdd(&m.lcnt[dig[e]], 1u) : 0u;
    u32 val[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) val[e] = LATE ? (ok[e] ? late_src[late_q + (u32)e * blockDim.x] : 0u) : val_in[e];
    __syncthreads();
    PK_PROF_ADD(prof_base + 1);
    const u32 total = block_exclusive_scan(m.lcnt, m.lstart, R, m.scratch);
    PK_PROF_ADD(prof_base + 2);
    for (u32 d = threadIdx.x; d < R; d += blockDim.x) {
        const u32 c = m.lcnt[d];
        if (c) m.gdelta[d] = atomicAdd(&gcursor[d], c) - m.lstart[d];
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        if (ok[e]) {
            const u32 pos = m.lstart[dig[e]] + rank[e];
            reinterpret_cast<uint2 *>(m.stage)[pos] = make_uint2(val[e], AUX ? ((dig[e] << 16) | aux[e]) : (dig[e] << 16));
        }
    }
    __syncthreads();
    PK_PROF_ADD(prof_base + 3);
    for (u32 q = threadIdx.x; q < total; q += blockDim.x) {
        const uint2 v = reinterpret_cast<const uint2 *>(m.stage)[q];
        const u32 kb = v.y;
        const u32 g = m.gdelta[kb >> 16] + q;
        out_val[g] = v.x;
        if (AUX) out_aux[g] = (u16)(kb & 0xffffu);
    }
    __syncthreads();
    PK_PROF_ADD(prof_base + 4);
    for (u32 d = threadIdx.x; d < R; d += blockDim.x) m.lcnt[d] = 0;
    __syncthreads();
    PK_PROF_ADD(prof_base + 5);
}

// Level 1: one block per tile of points, all windows.  Partitions (bucket, table index)
// pairs by the high bucket bits into l1_val (sign << 31 | table index) and l1_key (low bits).
#define PK_STAGE_EPT 16
__global__ void __launch_bounds__(1024) k_scatter_staged_b(const u32 *__restrict__ digits, MsmPlan p, u32 *__restrict__ gcursor,
                                                          u32 *__restrict__ l1_val, u16 *__restrict__ l1_key) {
    PK_DYN_SMEM(u32, smem);
    const u32 S = blockDim.x * PK_STAGE_EPT;
    const StageSmem m = stage_carve(smem, p.HI, S);
    for (u32 d = threadIdx.x; d < p.HI; d += blockDim.x) m.lcnt[d] = 0;
    __syncthreads();
    const u32 tile = blockIdx.x;
    const u32 beg = tile * p.tile;
    const u32 end = (beg + p.tile < p.n) ? beg + p.tile : p.n;
    const u32 lo_mask = (1u << p.lo_bits) - 1u;
    for (u32 w = 0; w < p.W; ++w) {
        const u32 *row = digits + (size_t)w * p.n_pad;
        const u32 base_idx = w * p.stride;
        for (u32 s0 = beg; s0 < end; s0 += S) {
            PK_PROF_START();
            u32 dig[PK_STAGE_EPT], val[PK_STAGE_EPT], aux[PK_STAGE_EPT];
            bool ok[PK_STAGE_EPT];
            if ((threadIdx.x & 7u) == 0) {  // the next stage's digits (same window further on, or the next window's row): one lane per 128-byte line
                const bool same = s0 + S < end;
                const u32 *nrow = same ? row : row + p.n_pad;
                const u32 ns0 = same ? s0 + S : beg;
                if (same || w + 1 < p.W) {
#pragma unroll
                    for (int u = 0; u < PK_STAGE_EPT / 4; ++u) {
                        const u32 i = ns0 + 4 * (threadIdx.x + (u32)u * blockDim.x);
                        if (i < end) prefetch_l2(nrow + i);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PK_STAGE_EPT / 4; ++u) {
                const u32 i = s0 + 4 * (threadIdx.x + (u32)u * blockDim.x);
                const uint4 v = (i < end) ? __ldg(reinterpret_cast<const uint4 *>(row + i)) : make_uint4(~0u, ~0u, ~0u, ~0u);
                const u32 d4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const u32 d = d4[k];
                    const u32 b = d & 0x7fffffffu;
                    ok[u * 4 + k] = (i + k < end) && d != 0xffffffffu;
                    dig[u * 4 + k] = b >> p.lo_bits;
                    aux[u * 4 + k] = b & lo_mask;
                    val[u * 4 + k] = (d & 0x80000000u) | (base_idx + i + (u32)k);
                }
            }
#ifdef PK_STAGE_PROF
            if (dig[0] == 0xdeadbeefu) l1_val[0] = 0;  // keep the loads ahead of the clock read
#endif
            PK_PROF_ADD(0);
            staged_partition<PK_STAGE_EPT>(p.HI, dig, val, aux, ok, m, gcursor, l1_val, l1_key);
        }
    }
}

// Level 1 fused with the decomposition (windows of 16 bits and more: W <= 16 entries per point): a stage is blockDim
// points, every thread converts ONE scalar (to_repr, fold to <= (r-1)/2, signed digits — the code of k_decompose_b)
// and contributes its W (bucket, table index) entries; all windows share the one bucket set, so they are partitioned
// together.  The digit array is never written.
template <int C>
__global__ void __launch_bounds__(1024) k_scatter_fused_b(const uint4 *__restrict__ scalars, MsmPlan p, u32 *__restrict__ gcursor,
                                                         u32 *__restrict__ l1_val, u16 *__restrict__ l1_key) {
    constexpr int W = (254 + C - 1) / C + ((C * ((254 + C - 1) / C) < 255) ? 1 : 0);
    constexpr u32 B = 1u << (C - 1);
    static_assert(W <= PK_STAGE_EPT, "the fused scatter holds one point's digits per thread");
    PK_DYN_SMEM(u32, smem);
    const u32 S = blockDim.x * PK_STAGE_EPT;
    const StageSmem m = stage_carve(smem, p.HI, S);
    for (u32 d = threadIdx.x; d < p.HI; d += blockDim.x) m.lcnt[d] = 0;
    __syncthreads();
    const u32 beg = blockIdx.x * p.tile;
    const u32 end = (beg + p.tile < p.n) ? beg + p.tile : p.n;
    const u32 lo_mask = (1u << p.lo_bits) - 1u;
    for (u32 s0 = beg; s0 < end; s0 += blockDim.x) {
        u32 dig[PK_STAGE_EPT], val[PK_STAGE_EPT], aux[PK_STAGE_EPT];
        bool ok[PK_STAGE_EPT];
#pragma unroll
        for (int e = 0; e < PK_STAGE_EPT; ++e) { ok[e] = false; dig[e] = 0; val[e] = 0; aux[e] = 0; }
        const u32 i = s0 + threadIdx.x;
        if (i < end) {
            fe v = fr_to_canonical(load_fe(scalars + 2 * (size_t)i));
            const bool neg = fr_above_half(v);
            if (neg) {
                u32 mm[8];
                FrMod::limbs(mm);
                fe t;
                sub8(t.l, mm, v.l);
                v = t;
            }
            u32 carry = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int off = w * C;
                const int word = off >> 5, sh = off & 31;
                u32 raw = 0;
                if (word < 8) {
                    raw = v.l[word] >> sh;
                    if (sh + C > 32 && word + 1 < 8) raw |= v.l[word + 1] << (32 - sh);
                }
                raw = (raw & ((1u << C) - 1u)) + carry;
                const bool borrow = neg ? (raw >= B) : (raw > B);  // see k_decompose
                const u32 mag = borrow ? (1u << C) - raw : raw;
                carry = borrow ? 1u : 0u;
                const u32 sign = (borrow ? 1u : 0u) ^ (neg ? 1u : 0u);
                if (mag != 0) {
                    const u32 b = mag - 1u;
                    ok[w] = true;
                    dig[w] = b >> p.lo_bits;
                    aux[w] = b & lo_mask;
                    val[w] = (sign << 31) | ((u32)w * p.stride + i);
                }
            }
        }
        staged_partition<PK_STAGE_EPT>(p.HI, dig, val, aux, ok, m, gcursor, l1_val, l1_key);
    }
}

// Level 2 works on slices of <= `slice` consecutive entries of one bin.  The slices of all bins are
// numbered consecutively (slice_prefix[bin] = number of the bin's first slice) and block b takes
// slice b, so a bin that holds far more than its share (a short top window, skewed scalars) is
// spread over as many blocks as it has slices.
__global__ void __launch_bounds__(1024) k_slice_prefix(const u32 *__restrict__ bin_start, u32 nbins, u32 slice, u32 *__restrict__ slice_prefix) {
    __shared__ u32 cnt[1024];
    __shared__ u32 start[1024];
    __shared__ u32 scratch[64];
    for (u32 b = threadIdx.x; b < 1024; b += blockDim.x) cnt[b] = (b < nbins) ? (bin_start[b + 1] - bin_start[b] + slice - 1) / slice : 0u;
    __syncthreads();
    const u32 total = block_exclusive_scan(cnt, start, 1024, scratch);
    for (u32 b = threadIdx.x; b < nbins; b += blockDim.x) slice_prefix[b] = start[b];
    if (threadIdx.x == 0) slice_prefix[nbins] = total;
}
// The (bin, entry range) of slice number b; false when b is past the last slice.
PK_HD bool slice_of_block(const u32 *__restrict__ slice_prefix, const u32 *__restrict__ bin_start, u32 nbins, u32 slice, u32 b, u32 &bin, u32 &sb,
                          u32 &se) {
    if (b >= slice_prefix[nbins]) return false;
    u32 lo = 0, hi = nbins;  // last bin with slice_prefix[bin] <= b (empty bins repeat the value: take the last)
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (slice_prefix[mid] <= b) lo = mid; else hi = mid;
    }
    bin = lo;
    const u32 end = bin_start[bin + 1];
    sb = bin_start[bin] + (b - slice_prefix[bin]) * slice;
    se = (sb + slice < end) ? sb + slice : end;
    return true;
}

// Level 2 histogram: block b counts the low bucket bits of slice b and adds them to the global
// per-bucket counts.
__global__ void __launch_bounds__(256) k_bucket_hist_b(const u16 *__restrict__ l1_key, MsmPlan p, const u32 *__restrict__ bin_start,
                                                       const u32 *__restrict__ slice_prefix, u32 slice, u32 *__restrict__ bucket_cnt) {
    __shared__ u32 cnt[4096];
    u32 bin, sb, se;
    if (!slice_of_block(slice_prefix, bin_start, p.nbins, slice, blockIdx.x, bin, sb, se)) return;
    const u32 nlo = 1u << p.lo_bits;
    for (u32 k = threadIdx.x; k < nlo; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    constexpr int U = 8;
    for (u32 q0 = sb + threadIdx.x; q0 < se; q0 += blockDim.x * U) {
        u32 k[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u32 q = q0 + (u32)u * blockDim.x;
            k[u] = (q < se) ? (u32)l1_key[q] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (k[u] != 0xffffffffu) atomicAdd(&cnt[k[u]], 1u);
        }
    }
    __syncthreads();
    for (u32 k = threadIdx.x; k < nlo; k += blockDim.x) {
        if (cnt[k]) atomicAdd(&bucket_cnt[(size_t)bin * nlo + k], cnt[k]);
    }
}

// In-place exclusive scan of v[0..n) (one block of 1024); copy[0..n] receives the same
// offsets plus the total at copy[n].  v then serves as the level's cursor array.
__global__ void __launch_bounds__(1024) k_scan_inplace(u32 *__restrict__ v, u32 n, u32 *__restrict__ copy) {
    __shared__ u32 scratch[64];
    const u32 per = (n + blockDim.x - 1) / blockDim.x;
    const u32 first = threadIdx.x * per;
    u32 sum = 0;
    for (u32 k = 0; k < per; ++k) {
        if (first + k < n) sum += v[first + k];
    }
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    u32 incl = sum;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 w = scratch[lane];
        u32 wi = w;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        scratch[lane] = wi - w;
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    u32 run = scratch[wid] + incl - sum;
    for (u32 k = 0; k < per; ++k) {
        if (first + k < n) {
            const u32 c = v[first + k];
            v[first + k] = run;
            copy[first + k] = run;
            run += c;
        }
    }
    if (threadIdx.x == 0) copy[n] = scratch[32];
}

// Multi-block version for large arrays: per-block totals, a one-block scan of the totals,
// then every block rescans its 2048 elements with its offset.
#define PK_SCAN_CHUNK 2048
__global__ void __launch_bounds__(1024) k_scan_partial(const u32 *__restrict__ v, u32 n, u32 *__restrict__ sums) {
    __shared__ u32 scratch[64];
    const u32 base = blockIdx.x * PK_SCAN_CHUNK;
    const u32 i0 = base + threadIdx.x, i1 = i0 + 1024;
    u32 sum = ((i0 < n) ? v[i0] : 0) + ((i1 < n) ? v[i1] : 0);
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (u32 d = 16; d >= 1; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
    if (lane == 0) scratch[wid] = sum;
    __syncthreads();
    if (wid == 0) {
        u32 w = (lane < (blockDim.x >> 5)) ? scratch[lane] : 0;
#pragma unroll
        for (u32 d = 16; d >= 1; d >>= 1) w += __shfl_down_sync(0xffffffffu, w, d);
        if (lane == 0) sums[blockIdx.x] = w;
    }
}
__global__ void __launch_bounds__(1024) k_scan_apply(u32 *__restrict__ v, u32 n, const u32 *__restrict__ offsets, u32 *__restrict__ copy) {
    __shared__ u32 cnt[PK_SCAN_CHUNK];
    __shared__ u32 start[PK_SCAN_CHUNK];
    __shared__ u32 scratch[64];
    const u32 base = blockIdx.x * PK_SCAN_CHUNK;
    for (u32 k = threadIdx.x; k < PK_SCAN_CHUNK; k += blockDim.x) cnt[k] = (base + k < n) ? v[base + k] : 0;
    __syncthreads();
    const u32 total = block_exclusive_scan(cnt, start, PK_SCAN_CHUNK, scratch);
    const u32 off = offsets[blockIdx.x];
    for (u32 k = threadIdx.x; k < PK_SCAN_CHUNK; k += blockDim.x) {
        if (base + k < n) {
            v[base + k] = off + start[k];
            copy[base + k] = off + start[k];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) copy[n] = off + total;
}

// Level 2 scatter: same slicing as the histogram; partitions each slice by the low bucket
// bits into the final (window-free) order.  gcursor = per-bucket cursors (bucket starts).
__global__ void __launch_bounds__(1024) k_bucket_scatter_staged_b(const u32 *__restrict__ l1_val, const u16 *__restrict__ l1_key, MsmPlan p,
                                                                 const u32 *__restrict__ bin_start, const u32 *__restrict__ slice_prefix, u32 slice,
                                                                 u32 *__restrict__ gcursor, u32 *__restrict__ sorted) {
    PK_DYN_SMEM(u32, smem);
    u32 bin, sb, se;
    if (!slice_of_block(slice_prefix, bin_start, p.nbins, slice, blockIdx.x, bin, sb, se)) return;
    const u32 nlo = 1u << p.lo_bits;
    const u32 S = blockDim.x * PK_STAGE_EPT;
    const StageSmem m = stage_carve(smem, nlo, S);
    for (u32 d = threadIdx.x; d < nlo; d += blockDim.x) m.lcnt[d] = 0;
    __syncthreads();
    u32 *cur = gcursor + (size_t)bin * nlo;
    for (u32 s0 = sb; s0 < se; s0 += S) {
        PK_PROF_START();
        u32 dig[PK_STAGE_EPT], val[PK_STAGE_EPT], aux[PK_STAGE_EPT];
        bool ok[PK_STAGE_EPT];
#pragma unroll
        for (int e = 0; e < PK_STAGE_EPT; ++e) {
            const u32 q = s0 + threadIdx.x + (u32)e * blockDim.x;
            ok[e] = q < se;
            dig[e] = ok[e] ? (u32)l1_key[q] : 0u;
            val[e] = 0;
            aux[e] = 0;
        }
#ifdef PK_STAGE_PROF
        if (dig[0] == 0xdeadbeefu) sorted[0] = 0;
#endif
        PK_PROF_ADD(8);
        staged_partition<PK_STAGE_EPT, false, true>(nlo, dig, val, aux, ok, m, cur, sorted, (u16 *)nullptr, 8, l1_val, s0 + threadIdx.x);
    }
}

// ---- building the table (once per registered base slice)
__global__ void __launch_bounds__(128) k_table_init(const affine *__restrict__ bases, u32 n, xyzz *__restrict__ cur, affine *__restrict__ row0) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *bp = reinterpret_cast<const uint4 *>(bases + i);
    affine a;
    a.x = load_fe(bp);
    a.y = load_fe(bp + 2);
    uint4 *q = reinterpret_cast<uint4 *>(row0 + i);
    store_fe(q, a.x);
    store_fe(q + 2, a.y);
    store_xyzz(cur + i, xyzz_from_affine(a));
}
__global__ void __launch_bounds__(128) k_table_double(xyzz *__restrict__ cur, u32 n, u32 c) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz v = load_xyzz(cur + i);
    for (u32 k = 0; k < c; ++k) v = xyzz_double(v);
    store_xyzz(cur + i, v);
}
// Thread normalises 8 consecutive points with one inversion (Montgomery's trick).
__global__ void __launch_bounds__(128) k_table_normalize(const xyzz *__restrict__ cur, u32 n, affine *__restrict__ row) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 i0 = t * 8;
    if (i0 >= n) return;
    const u32 cnt = (n - i0 < 8) ? n - i0 : 8;
    fe prefix[8];
    fe acc = fq_one();
    for (u32 j = 0; j < cnt; ++j) {
        const xyzz v = load_xyzz(cur + i0 + j);
        prefix[j] = acc;
        if (!xyzz_is_identity(v)) acc = fq_mul(acc, fq_mul(v.zz, v.zzz));
    }
    fe inv = fq_inv_fast(acc);
    for (u32 jj = cnt; jj-- > 0;) {
        const xyzz v = load_xyzz(cur + i0 + jj);
        affine r;
        if (xyzz_is_identity(v)) {
            r.x = fe_zero(); r.y = fe_zero();
        } else {
            const fe i = fq_mul(inv, prefix[jj]);  // 1 / (zz * zzz)
            inv = fq_mul(inv, fq_mul(v.zz, v.zzz));
            r.x = fq_mul(v.x, fq_mul(i, v.zzz));
            r.y = fq_mul(v.y, fq_mul(i, v.zz));
        }
        uint4 *q = reinterpret_cast<uint4 *>(row + i0 + jj);
        store_fe(q, r.x);
        store_fe(q + 2, r.y);
    }
}

// ==================================================== warp segmented reduction
PK_HD u32 shfl_u32(u32 v, int src) { return __shfl_sync(0xffffffffu, v, src); }

#ifndef PLONKISH_EMUL
PK_HD xyzz shfl_down_xyzz(const xyzz &v, u32 d) {
    xyzz r;
    const u32 *src = reinterpret_cast<const u32 *>(&v);
    u32 *dst = reinterpret_cast<u32 *>(&r);
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i] = __shfl_down_sync(0xffffffffu, src[i], d);
    return r;
}
PK_HD xyzz shfl_up_xyzz(const xyzz &v, u32 d) {
    xyzz r;
    const u32 *src = reinterpret_cast<const u32 *>(&v);
    u32 *dst = reinterpret_cast<u32 *>(&r);
#pragma unroll
    for (int i = 0; i < 32; ++i) dst[i] = __shfl_up_sync(0xffffffffu, src[i], d);
    return r;
}
#else  // emulation moves the whole point in one exchange (32x fewer barriers)
inline xyzz shfl_down_xyzz(const xyzz &v, u32 d) { return emul::warp_exchange_blob(v, (int)d); }
inline xyzz shfl_up_xyzz(const xyzz &v, u32 d) { return emul::warp_exchange_blob(v, -(int)d); }
#endif

// Every lane brings the two open ends of its run: head (hk, hp) = the sum for the
// first key it saw, tail (tk, tp) = the sum for the last key if different (else
// tk == hk and tp = identity).  Keys are non-decreasing across lanes and within a
// lane.  Sums that are provably complete inside this warp are stored to
// bucket_sum[key]; the warp's first and last open sums go to the next level as
// items 2*warp and 2*warp+1 (terminal: everything is stored).
PK_HD void warp_merge(u32 hk, xyzz hp, u32 tk, const xyzz &tp, xyzz *bucket_sum, u32 *out_keys, xyzz *out_pts,
                      u32 warp_global, bool terminal) {
    const u32 lane = threadIdx.x & 31u;
    const bool real_tail = (tk != hk);
    // A: a lane's tail continues in the next lane's head.
    const u32 prev_tk = __shfl_up_sync(0xffffffffu, tk, 1);
    const xyzz prev_tp = shfl_up_xyzz(tp, 1);
    const u32 next_hk = __shfl_down_sync(0xffffffffu, hk, 1);
    if (lane > 0 && prev_tk == hk) hp = xyzz_add(hp, prev_tp);
    const bool tail_consumed = (lane < 31) && (next_hk == tk);
    if (real_tail && lane < 31 && !tail_consumed && tk != PK_INVALID_KEY) store_xyzz(bucket_sum + tk, tp);
    // B: segmented reduction of the heads.
#pragma unroll 1
    for (u32 d = 1; d < 32; d <<= 1) {
        const u32 okey = __shfl_down_sync(0xffffffffu, hk, d);
        const xyzz other = shfl_down_xyzz(hp, d);
        if (lane + d < 32 && okey == hk) hp = xyzz_add(hp, other);
    }
    const u32 prev_hk = __shfl_up_sync(0xffffffffu, hk, 1);
    const bool leader = (lane == 0) || (prev_hk != hk);
    const u32 hk31 = shfl_u32(hk, 31), tk31 = shfl_u32(tk, 31), hk0 = shfl_u32(hk, 0);
    const bool rt31 = (tk31 != hk31);
    if (terminal) {
        if (leader && hk != PK_INVALID_KEY) store_xyzz(bucket_sum + hk, hp);
        if (lane == 31 && real_tail && tk != PK_INVALID_KEY) store_xyzz(bucket_sum + tk, tp);
        return;
    }
    if (leader) {
        if (lane == 0) {
            out_keys[2 * warp_global] = hk;
            store_xyzz(out_pts + 2 * warp_global, hp);
        } else if (!rt31 && hk == hk31) {
            out_keys[2 * warp_global + 1] = hk;
            store_xyzz(out_pts + 2 * warp_global + 1, hp);
        } else if (hk != PK_INVALID_KEY) {
            store_xyzz(bucket_sum + hk, hp);
        }
    }
    if (rt31) {
        if (lane == 31) {
            out_keys[2 * warp_global + 1] = tk;
            store_xyzz(out_pts + 2 * warp_global + 1, tp);
        }
    } else if (hk0 == hk31) {
        if (lane == 0) {
            out_keys[2 * warp_global + 1] = hk31;
            store_xyzz(out_pts + 2 * warp_global + 1, xyzz_identity());
        }
    }
}

// ================================================================ K3 accumulate
// Thread t sums sorted[t*L, (t+1)*L) and leaves two items for the segmented
// reduction: item 2t = (first key of the run, its sum), item 2t+1 = (last key, its
// sum), the second being (same key, identity) when the run stays inside one bucket.
// Buckets that begin and end inside the run are complete and stored directly.
// bucket_start[nbuckets] is the entry count.  The kernel keeps only the running
// XYZZ sum and one affine point live so that four warps fit per SM sub-partition.
#ifndef PK_ACC_MIN_BLOCKS
#define PK_ACC_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, PK_ACC_MIN_BLOCKS) k_accumulate(const u32 *__restrict__ sorted, const u32 *__restrict__ bucket_start,
                                                       const affine *__restrict__ bases, MsmPlan p, xyzz *__restrict__ bucket_sum,
                                                       u32 *__restrict__ out_keys, xyzz *__restrict__ out_pts) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 total = bucket_start[p.nbuckets];
    u32 hk = PK_INVALID_KEY, tk = PK_INVALID_KEY;
    xyzz acc = xyzz_identity();
    bool head_written = false;
    u32 s = 0, e = 0;
    if (pk_acc_run(t, total, p.acc_tiers, p.acc_resident, p.nthreads1, p.L, s, e)) {
        // largest g with bucket_start[g] <= s
        u32 lo = 0, hi = p.nbuckets;
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if (bucket_start[mid] <= s) lo = mid; else hi = mid;
        }
        u32 g = lo;
        u32 next = bucket_start[g + 1];
        hk = g;
        for (u32 pos = s; pos < e; ++pos) {
            if (pos == next) {
                acc = xyzz_canonical(acc);  // the loop keeps coordinates in [0, 2p)
                if (!head_written) { store_xyzz(out_pts + 2 * (size_t)t, acc); head_written = true; } else { store_xyzz(bucket_sum + g, acc); }
                acc = xyzz_identity();
                do { ++g; next = bucket_start[g + 1]; } while (next <= pos);
            }
            const u32 entry = sorted[pos];
            const uint4 *bp = reinterpret_cast<const uint4 *>(bases + (entry & 0x7fffffffu));
            const fe x = load_fe(bp);
            fe y = load_fe(bp + 2);
            if (entry >> 31) y = fq_neg(y);
            xyzz_madd_lazy(acc, x, y);
        }
        tk = g;
        acc = xyzz_canonical(acc);
    }
    out_keys[2 * (size_t)t] = hk;
    out_keys[2 * (size_t)t + 1] = tk;
    if (!head_written) {
        store_xyzz(out_pts + 2 * (size_t)t, acc);
        acc = xyzz_identity();
    }
    store_xyzz(out_pts + 2 * (size_t)t + 1, acc);
}

// ======================================================= K3b item reduce levels
// Lane handles K consecutive (key, point) items of the previous level.
__global__ void __launch_bounds__(128) k_reduce_items(const u32 *__restrict__ in_keys, const xyzz *__restrict__ in_pts, u32 count,
                                                      u32 K, xyzz *__restrict__ bucket_sum, u32 *__restrict__ out_keys,
                                                      xyzz *__restrict__ out_pts, int terminal) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long s64 = (unsigned long long)t * K;
    u32 hk = PK_INVALID_KEY, tk = PK_INVALID_KEY;
    xyzz hp = xyzz_identity(), acc = xyzz_identity();
    if (s64 < count) {
        const u32 s = (u32)s64;
        const u32 e = (s + K < count) ? s + K : count;
        u32 g = in_keys[s];
        acc = load_xyzz(in_pts + s);
        u32 nflushed = 0;
        for (u32 pos = s + 1; pos < e; ++pos) {
            const u32 k = in_keys[pos];
            const xyzz v = load_xyzz(in_pts + pos);
            if (k != g) {
                if (nflushed == 0) { hk = g; hp = acc; } else if (g != PK_INVALID_KEY) { store_xyzz(bucket_sum + g, acc); }
                ++nflushed;
                g = k;
                acc = v;
            } else {
                acc = xyzz_add(acc, v);
            }
        }
        if (nflushed == 0) { hk = g; hp = acc; tk = g; acc = xyzz_identity(); } else { tk = g; }
    }
    warp_merge(hk, hp, tk, acc, bucket_sum, out_keys, out_pts, t >> 5, terminal != 0);
}

// ============================================================== K4 bucket reduce
PK_HD xyzz warp_sum_xyzz(xyzz v) {
#pragma unroll 1
    for (u32 d = 16; d >= 1; d >>= 1) {
        const xyzz o = shfl_down_xyzz(v, d);
        v = xyzz_add(v, o);
    }
    return v;  // lane 0 holds the sum
}

// k * P by double-and-add over the bits of k.
PK_HD xyzz xyzz_mul_small(const xyzz &pnt, u32 k) {
    xyzz r = xyzz_identity();
    for (int b = 31 - __clz(k | 1u); b >= 0; --b) {
        r = xyzz_double(r);
        if ((k >> b) & 1u) r = xyzz_add(r, pnt);
    }
    return r;
}

// grid (red_blocks, W), block 256.  Thread j of window w owns buckets
// [j*rb, (j+1)*rb): sum_i (j*rb + i + 1) * B_i = acc + (j*rb) * run, where run is
// the plain sum and acc the running-sum total (msm.rs:175-179 restated per chunk).
__global__ void __launch_bounds__(PK_RED_BLOCK, 3) k_bucket_reduce(const xyzz *__restrict__ bucket_sum,
                                                                MsmPlan p, xyzz *__restrict__ block_out) {
    __shared__ xyzz warp_part[PK_RED_BLOCK / 32];
    const u32 w = blockIdx.y;
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    xyzz contrib = xyzz_identity();
    if (j < p.red_threads) {
        xyzz run = xyzz_identity(), acc = xyzz_identity();
        const u32 first = j * p.rb;
        const u32 cnt = (first + p.rb <= p.B) ? p.rb : p.B - first;  // the last thread's range may be short
        for (int i = (int)cnt - 1; i >= 0; --i) {
            const u32 g = w * p.B + first + (u32)i;
            for (u32 k = 0; k < p.nchunks; ++k)  // untouched buckets are zeroed = identity
                run = xyzz_add(run, load_xyzz(bucket_sum + (size_t)k * p.nbuckets + g));
            acc = xyzz_add(acc, run);
        }
        contrib = xyzz_add(acc, xyzz_mul_small(run, j * p.rb));
    }
    contrib = warp_sum_xyzz(contrib);
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if (lane == 0) warp_part[wid] = contrib;
    __syncthreads();
    if (wid == 0) {
        xyzz v = (lane < (blockDim.x >> 5)) ? warp_part[lane] : xyzz_identity();
        v = warp_sum_xyzz(v);
        if (lane == 0) store_xyzz(block_out + (size_t)w * p.red_blocks + blockIdx.x, v);
    }
}

// ============================================================ K5 window combine
// One warp per window: S_w = sum of the window's block partials, T_w = 2^(c*w) * S_w.
__global__ void __launch_bounds__(32) k_window_weight(const xyzz *__restrict__ block_out, MsmPlan p, xyzz *__restrict__ win_out) {
    const u32 lane = threadIdx.x & 31u, w = blockIdx.x;
    xyzz v = xyzz_identity();
    for (u32 k = lane; k < p.red_blocks; k += 32) v = xyzz_add(v, load_xyzz(block_out + (size_t)w * p.red_blocks + k));
    v = warp_sum_xyzz(v);
    if (lane == 0) {
        for (u32 k = 0; k < p.c * w; ++k) v = xyzz_double(v);
        store_xyzz(win_out + w, v);
    }
}
// result = sum_w T_w (+ *prev: the running total of earlier chunks of the same MSM).
__global__ void __launch_bounds__(32) k_window_sum(const xyzz *__restrict__ win_out, u32 W, const xyzz *prev, xyzz *__restrict__ result) {
    const u32 lane = threadIdx.x & 31u;
    xyzz t = (lane < W) ? load_xyzz(win_out + lane) : xyzz_identity();
    t = warp_sum_xyzz(t);
    if (lane == 0) {
        if (prev) t = xyzz_add(t, load_xyzz(prev));
        store_xyzz(result, t);
    }
}

// One bucket set (mode 1): the window combine is just the sum of the reduce blocks' partials (+ *prev).  One block of
// 256 threads instead of the one-warp k_window_weight + k_window_sum pair: 2 + 5 + 3 dependent additions for ~300
// partials instead of 10 + 5 + 5 and a launch (0.126 -> 0.06 ms per MSM; matters for small MSMs and sharded slices).
__global__ void __launch_bounds__(256) k_combine_single(const xyzz *__restrict__ block_out, u32 count, const xyzz *prev, xyzz *__restrict__ result) {
    __shared__ xyzz warp_part[8];
    const u32 lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    xyzz v = xyzz_identity();
    for (u32 k = threadIdx.x; k < count; k += blockDim.x) v = xyzz_add(v, load_xyzz(block_out + k));
    v = warp_sum_xyzz(v);
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    if (wid == 0) {
        xyzz t = (lane < (blockDim.x >> 5)) ? warp_part[lane] : xyzz_identity();
#pragma unroll 1
        for (u32 d = 4; d >= 1; d >>= 1) {
            const xyzz o = shfl_down_xyzz(t, d);
            t = xyzz_add(t, o);
        }
        if (lane == 0) {
            if (prev) t = xyzz_add(t, load_xyzz(prev));
            store_xyzz(result, t);
        }
    }
}

// Sum `count` projective partials (one per chunk or per GPU) and normalise:
// the `.to_affine()` every reference caller applies (e.g. kzg.rs:255, pcs.rs:175).
__global__ void k_finalize(const xyzz *__restrict__ partials, u32 count, affine *__restrict__ out_affine, xyzz *__restrict__ out_xyzz) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    xyzz t = xyzz_identity();
    for (u32 k = 0; k < count; ++k) t = xyzz_add(t, load_xyzz(partials + k));
    if (out_xyzz) store_xyzz(out_xyzz, t);
    if (out_affine) {
        const affine a = xyzz_to_affine(t);
        uint4 *q = reinterpret_cast<uint4 *>(out_affine);
        store_fe(q, a.x);
        store_fe(q + 2, a.y);
    }
}

// ================================================================ launch sequence
template <int C>
inline void pk_launch_decompose(const MsmPlan &p, const void *scalars, const MsmWorkspace &ws, pk_stream_t stream) {
    const size_t smem = sizeof(u32) * p.nbins;
#ifndef PLONKISH_EMUL
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_decompose<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    PK_LAUNCH(k_decompose<C>, dim3(p.ntiles), dim3(p.blk), smem, stream, (const uint4 *)scalars, p, ws.digits, ws.tile_hist);
}

// Optional stage boundaries for the profiling entry point: ev[0] before K1, then one
// event after decompose, scans, scatter, sort, accumulate, item levels, bucket
// reduce, window combine (9 events).
struct StageMarks {
    void *ev[9];
};

// In-place exclusive scan of v[0..n) with a copy (+ total at copy[n]); multi-block above 4096.
inline void pk_enqueue_scan(u32 *v, u32 n, u32 *copy, u32 *scan_tmp, pk_stream_t stream) {
    if (n <= 2 * PK_SCAN_CHUNK) {
        PK_LAUNCH(k_scan_inplace, dim3(1), dim3(1024), 0, stream, v, n, copy);
    } else {
        const u32 nb = (n + PK_SCAN_CHUNK - 1) / PK_SCAN_CHUNK;  // <= 2048 for c <= 22
        PK_LAUNCH(k_scan_partial, dim3(nb), dim3(1024), 0, stream, v, n, scan_tmp);
        PK_LAUNCH(k_scan_inplace, dim3(1), dim3(1024), 0, stream, scan_tmp, nb, scan_tmp + 2050);
        PK_LAUNCH(k_scan_apply, dim3(nb), dim3(1024), 0, stream, v, n, scan_tmp, copy);
    }
}

// Phase 1 of an MSM over p.n points: decompose, sort, accumulate into bucket array p.chunk of
// ws.bucket_sum (zeroed first).  Chunks of one MSM share c and the bucket layout; the
// reduce phase adds their arrays.
inline void pk_enqueue_buckets(const MsmPlan &p, const void *scalars, const void *bases, const MsmWorkspace &ws, pk_stream_t stream,
                               const StageMarks *marks = nullptr) {
    PK_MARK(marks, 0, stream);
    xyzz *bsum = ws.bucket_sum + (size_t)p.chunk * p.nbuckets;
    PK_MEMSET0(bsum, sizeof(xyzz) * (size_t)p.nbuckets, stream);
    if (p.mode == 0) {
        switch (p.c) {
            case 8: pk_launch_decompose<8>(p, scalars, ws, stream); break;
            case 9: pk_launch_decompose<9>(p, scalars, ws, stream); break;
            case 10: pk_launch_decompose<10>(p, scalars, ws, stream); break;
            case 11: pk_launch_decompose<11>(p, scalars, ws, stream); break;
            case 12: pk_launch_decompose<12>(p, scalars, ws, stream); break;
            case 13: pk_launch_decompose<13>(p, scalars, ws, stream); break;
            case 14: pk_launch_decompose<14>(p, scalars, ws, stream); break;
            case 15: pk_launch_decompose<15>(p, scalars, ws, stream); break;
            default: pk_launch_decompose<16>(p, scalars, ws, stream); break;
        }
        PK_MARK(marks, 1, stream);
        PK_LAUNCH(k_scan_tiles, dim3((p.nbins + p.blk - 1) / p.blk), dim3(p.blk), 0, stream, ws.tile_hist, p.ntiles, p.nbins, ws.bin_total);
        PK_LAUNCH(k_scan_bins, dim3(1), dim3(1024), 0, stream, ws.bin_total, p.nbins, ws.bin_start);
        PK_MARK(marks, 2, stream);
        PK_LAUNCH(k_scatter_bins, dim3(p.ntiles, p.W), dim3(p.blk), 0, stream, ws.digits, p, ws.tile_hist, ws.bin_start, ws.l1);
        PK_MARK(marks, 3, stream);
        PK_LAUNCH(k_sort_bins, dim3(p.nbins), dim3(p.blk), 0, stream, ws.l1, p, ws.bin_start, ws.sorted, ws.bucket_start);
        PK_MARK(marks, 4, stream);
    } else {
        u32 *digits32 = reinterpret_cast<u32 *>(ws.digits);
        const size_t emax = (size_t)p.n * p.W;
        u32 *l1_val = ws.l1;
        u16 *l1_key = reinterpret_cast<u16 *>(ws.l1 + emax);
        PK_MEMSET0(ws.bin_total, sizeof(u32) * p.HI, stream);
        PK_MEMSET0(ws.bucket_cur, sizeof(u32) * (p.nbuckets + 1), stream);
        const bool fused = p.c >= 16 && p.fuse_l1;  // W <= 16: level 1 recomputes the digits, the first pass only counts
#define PK_DECOMPOSE_B(C) case C: PK_LAUNCH(k_decompose_b<C>, dim3(p.ntiles), dim3(p.blk), 0, stream, (const uint4 *)scalars, p, digits32, ws.bin_total); break;
#define PK_COUNT_B(C) case C: PK_LAUNCH((k_decompose_b<C, false>), dim3(p.ntiles), dim3(p.blk), 0, stream, (const uint4 *)scalars, p, digits32, ws.bin_total); break;
        if (fused) {
            switch (p.c) {
                PK_COUNT_B(16) PK_COUNT_B(17) PK_COUNT_B(18) PK_COUNT_B(19) PK_COUNT_B(20) PK_COUNT_B(21)
                default: PK_LAUNCH((k_decompose_b<22, false>), dim3(p.ntiles), dim3(p.blk), 0, stream, (const uint4 *)scalars, p, digits32, ws.bin_total); break;
            }
        } else {
            switch (p.c) {
                PK_DECOMPOSE_B(8) PK_DECOMPOSE_B(9) PK_DECOMPOSE_B(10) PK_DECOMPOSE_B(11) PK_DECOMPOSE_B(12)
                PK_DECOMPOSE_B(13) PK_DECOMPOSE_B(14) PK_DECOMPOSE_B(15) PK_DECOMPOSE_B(16) PK_DECOMPOSE_B(17)
                PK_DECOMPOSE_B(18) PK_DECOMPOSE_B(19) PK_DECOMPOSE_B(20) PK_DECOMPOSE_B(21)
                default: PK_LAUNCH(k_decompose_b<22>, dim3(p.ntiles), dim3(p.blk), 0, stream, (const uint4 *)scalars, p, digits32, ws.bin_total); break;
            }
        }
#undef PK_DECOMPOSE_B
#undef PK_COUNT_B
        PK_MARK(marks, 1, stream);
        // bin_total becomes the level-1 cursor array, bin_start the bin offsets (+ total).
        PK_LAUNCH(k_scan_inplace, dim3(1), dim3(1024), 0, stream, ws.bin_total, p.HI, ws.bin_start);
        PK_MARK(marks, 2, stream);
        const u32 S = p.blk_stage * PK_STAGE_EPT;
        const size_t smem1 = stage_smem_bytes(p.HI, S);
        if (fused) {
#define PK_FUSED_B(C) case C: PK_SET_SMEM(k_scatter_fused_b<C>, smem1); PK_LAUNCH(k_scatter_fused_b<C>, dim3(p.ntiles), dim3(p.blk_stage), smem1, stream, (const uint4 *)scalars, p, ws.bin_total, l1_val, l1_key); break;
            switch (p.c) {
                PK_FUSED_B(16) PK_FUSED_B(17) PK_FUSED_B(18) PK_FUSED_B(19) PK_FUSED_B(20) PK_FUSED_B(21)
                default: PK_SET_SMEM(k_scatter_fused_b<22>, smem1); PK_LAUNCH(k_scatter_fused_b<22>, dim3(p.ntiles), dim3(p.blk_stage), smem1, stream, (const uint4 *)scalars, p, ws.bin_total, l1_val, l1_key); break;
            }
#undef PK_FUSED_B
        } else {
            PK_SET_SMEM(k_scatter_staged_b, smem1);
            PK_LAUNCH(k_scatter_staged_b, dim3(p.ntiles), dim3(p.blk_stage), smem1, stream, digits32, p, ws.bin_total, l1_val, l1_key);
        }
        PK_MARK(marks, 3, stream);
        const u32 slice = 4 * S;
        const u32 max_slices = (u32)((emax + slice - 1) / slice) + p.nbins;  // every bin may end in a partial slice
        PK_LAUNCH(k_slice_prefix, dim3(1), dim3(1024), 0, stream, ws.bin_start, p.nbins, slice, ws.slice_prefix);
        PK_LAUNCH(k_bucket_hist_b, dim3(max_slices), dim3(p.blk), 0, stream, l1_key, p, ws.bin_start, ws.slice_prefix, slice, ws.bucket_cur);
        pk_enqueue_scan(ws.bucket_cur, p.nbuckets, ws.bucket_start, ws.scan_tmp, stream);
        const size_t smem2 = stage_smem_bytes(1u << p.lo_bits, S);
        PK_SET_SMEM(k_bucket_scatter_staged_b, smem2);
        PK_LAUNCH(k_bucket_scatter_staged_b, dim3(max_slices), dim3(p.blk_stage), smem2, stream, l1_val, l1_key, p, ws.bin_start, ws.slice_prefix,
                  slice, ws.bucket_cur, ws.sorted);
        PK_MARK(marks, 4, stream);
    }

    // K3, then the item levels until one warp stores everything that is left.
    PK_LAUNCH(k_accumulate, dim3(p.nthreads1 / p.blk_acc), dim3(p.blk_acc), 0, stream, ws.sorted, ws.bucket_start,
              (const affine *)bases, p, bsum, ws.item_keys[0], ws.item_pts[0]);
    PK_MARK(marks, 5, stream);
    u32 count = 2 * p.nthreads1;
    int src = 0;
    for (int terminal = 0; !terminal;) {
        const u32 K = (count > p.serial_items) ? 8 : 1;
        const u32 lanes = (count + K - 1) / K;
        const u32 nthreads = (lanes <= 32) ? 32u : ((lanes + 127u) & ~127u);
        const u32 warps = nthreads / 32;
        terminal = (warps == 1) ? 1 : 0;
        PK_LAUNCH(k_reduce_items, dim3(terminal ? 1 : nthreads / 128), dim3(terminal ? 32 : 128), 0, stream, ws.item_keys[src], ws.item_pts[src], count, K,
                  bsum, ws.item_keys[src ^ 1], ws.item_pts[src ^ 1], terminal);
        count = 2 * warps;
        src ^= 1;
    }
    PK_MARK(marks, 6, stream);
}

// Phase 2: bucket reduce and window combine; the projective result lands in ws.result
// (plus *prev if given).
inline void pk_enqueue_reduce(const MsmPlan &p, const MsmWorkspace &ws, const xyzz *prev, pk_stream_t stream,
                              const StageMarks *marks = nullptr) {
    PK_LAUNCH(k_bucket_reduce, dim3(p.red_blocks, p.ngroups), dim3(PK_RED_BLOCK), 0, stream, ws.bucket_sum, p, ws.block_out);
    PK_MARK(marks, 7, stream);
    if (p.ngroups == 1) {
        PK_LAUNCH(k_combine_single, dim3(1), dim3(256), 0, stream, ws.block_out, p.red_blocks, prev, ws.result);
    } else {
        PK_LAUNCH(k_window_weight, dim3(p.ngroups), dim3(32), 0, stream, ws.block_out, p, ws.win_out);
        PK_LAUNCH(k_window_sum, dim3(1), dim3(32), 0, stream, ws.win_out, p.ngroups, prev, ws.result);
    }
    PK_MARK(marks, 8, stream);
}

// One whole MSM.  No host synchronisation.
inline void pk_enqueue_msm(const MsmPlan &p, const void *scalars, const void *bases, const MsmWorkspace &ws, const xyzz *prev,
                           pk_stream_t stream, const StageMarks *marks = nullptr) {
    pk_enqueue_buckets(p, scalars, bases, ws, stream, marks);
    pk_enqueue_reduce(p, ws, prev, stream, marks);
}

// Builds the table of window multiples for n bases: table[w*n + i] = 2^(c*w) * bases[i],
// affine, W rows.  cur is n XYZZ points of scratch.
inline void pk_enqueue_table_build(const void *bases, u32 n, u32 c, u32 W, xyzz *cur, affine *table, pk_stream_t stream) {
    const u32 blocks = (n + 127) / 128;
    PK_LAUNCH(k_table_init, dim3(blocks), dim3(128), 0, stream, (const affine *)bases, n, cur, table);
    for (u32 w = 1; w < W; ++w) {
        PK_LAUNCH(k_table_double, dim3(blocks), dim3(128), 0, stream, cur, n, c);
        PK_LAUNCH(k_table_normalize, dim3(((n + 7) / 8 + 127) / 128), dim3(128), 0, stream, cur, n, table + (size_t)w * n);
    }
}

}  // namespace pk
