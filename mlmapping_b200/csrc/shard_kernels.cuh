// One logical map sharded over several GPUs ("large LiDAR scans"): stage 1 by phi column
// (rank r casts the columns phi % world == r: hit folds and miss bitmaps of different columns are
// disjoint), stage 2 by subbox owner (hash(glb) % world).  Between the stages the ranks exchange
//   * the distinct hit keys with their first-insert stamps (an all-gather) so that every rank derives the same
//     libstdc++ iteration order for the whole frame, and
//   * fixed-size update records per touched voxel (an all-to-all), consumed by the owner's k_fuse.
// Both exchanges are done BY THESE KERNELS over NVLink peer memory: every rank owns an exchange arena
// (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by its peers); a source writes its keys and records
// straight into the region reserved for it in the destination's arena, then publishes its counts and raises an
// epoch flag with a system-scope release; the destination's wait kernel acquires all flags before the owner-side
// kernels run.  No host round trip, no library collective and no remote atomic on the data path; the arenas are
// double-buffered by scan parity (a source can be at most one scan ahead of a destination, because its own wait
// needs the destination's flag of the scan before).  With world == 1 the same kernels run on one GPU.
#pragma once
#include "frame_kernels.cuh"
#include "order_kernels.cuh"

namespace mlm {

struct __align__(8) ShardRecord {  // 24 bytes
  int c[3];     // canonical global cell coordinate of the voxel
  int key;      // awareness cell index of a hit key, or -1 for a miss record
  float p;      // hit: folded probability
  int count;    // miss: number of miss cells mapping to the voxel; hit: first-insert stamp of the key
};

constexpr int kMaxWorld = 16;
// one rank's exchange arena as seen from this rank (peer-mapped addresses for the other ranks)
struct ShardArena {
  uint32_t *flags;        // [kMaxWorld] epoch of the last complete scan of every source
  int2 *mbox;             // [2][kMaxWorld] by parity, source: {distinct hit keys cast, records sent} (-1: the source failed)
  int *gather_key;        // [2][world][hit_cap] by parity, source: the source's distinct hit keys ...
  uint32_t *gather_stamp; // [2][world][hit_cap] ... and their first-insert stamps
  int *cursor;            // [2] by parity: records reserved in the inbox so far (sources reserve with a remote atomicAdd)
  ShardRecord *inbox;     // [2][rec_cap] by parity: update records for voxels this rank owns, all sources interleaved
  uint32_t *pflags;       // [kMaxWorld] epoch of the last scan whose point slice the source has delivered here
  double *points;         // [2][3 * max_points] by parity: the scan, assembled from the ranks' slices (submit_points_slice)
};
struct ShardPeers {
  ShardArena a[kMaxWorld];
  int rank, world;
  int hit_cap, rec_cap;
  int max_points;
};
// what the wait kernel found (device copy read by the owner-side kernels, pinned host copy read by the caller)
struct ShardState {
  int n_total;            // distinct hit keys of the scan over all ranks
  int n_rec_total;        // records received
  int rehash;             // 1: n_total exceeds the bucket count -> the owner-side kernels skipped, host runs the rehash path
  int error;              // device error code (peer failure, timeout, capacity)
  int hit_base, touched_base;  // entries of the local hit / touched lists in front of the ingested records
  int cnt_hits[kMaxWorld];
  unsigned long long wait_ns;  // time the wait kernel spent spinning (exchange skew seen by this rank)
};
constexpr int kErrPeer = 101;  // a peer did not signal in time / reported a failure

__device__ __forceinline__ int owner_of(const MapParams &P, const int c[3], int world, int g[3]) {
  g[0] = fast_floor_div(c[0], P.n, P.n_mul, P.n_shift);
  g[1] = fast_floor_div(c[1], P.n, P.n_mul, P.n_shift);
  g[2] = fast_floor_div(c[2], P.n, P.n_mul, P.n_shift);
  return subbox_owner(g, world);
}

// CTA-wide slot reservation: every thread asks for `want` (0 or more) consecutive slots of a global counter; one
// atomicAdd per CTA and call (a per-warp atomic on one address serialises: 33 k of them cost 130 us per CFG-C scan)
__device__ __forceinline__ int cta_reserve(int *ctr, int want, int *s_warp /*[33]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int incl = want;
#pragma unroll
  for (int ofs = 1; ofs < 32; ofs <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, ofs);
    if (lane >= ofs) incl += t;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int v = lane < nw ? s_warp[lane] : 0;
    int si = v;
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, si, ofs);
      if (lane >= ofs) si += t;
    }
    const int total = __shfl_sync(0xffffffffu, si, 31);
    int base = 0;
    if (lane == 0 && total > 0) base = atomicAdd(ctr, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < nw) s_warp[lane] = base + si - v;
  }
  __syncthreads();
  const int res = s_warp[w] + incl - want;
  __syncthreads();  // s_warp is reused by the next call
  return res;
}

// start of a scan: everything the staging and the push kernels expect to find reset, in one launch
//   act[0, B)        activation stamps the staging kernels atomicMin into (stale stamps of earlier scans must not survive)
//   scratch[...]     push ticket and skip flag
//   cursor_prev      the inbox cursor of the PREVIOUS scan's parity: its records are ingested, and no source can reserve in
//                    it again before it has seen this scan's flag (raised further down this stream)
__global__ void __launch_bounds__(256) k_shard_begin(uint32_t *act, uint32_t B, int *scratch, int n_scratch, int *cursor_prev) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (uint32_t i = gtid; i < B; i += nth) act[i] = 0xffffffffu;
  if (gtid < n_scratch) scratch[gtid] = 0;
  if (gtid == 0) *cursor_prev = 0;
}

constexpr int kPushThreads = 1024;  // list entries per chunk = per reservation in the destinations' inboxes
constexpr int kPushStage = 1792;   // records a CTA chunk stages in shared memory (42 KB)
// ---- source side: ONE kernel per scan pushes everything this rank has for the others ----------------------------
// (1) all-gather of the rank's distinct hit keys + stamps: written into every rank's gather region for this source;
// (2) all-to-all of the update records: one thread per entry of the list of voxels OTHER ranks own (the staging kept the
//     voxels this rank owns itself in the ordinary touched list, with their subboxes resolved: they never leave and meet
//     the other ranks' records in the owner-side ingest).  Records of a CTA chunk are grouped by
//     destination in shared memory, each group reserves its slots in the DESTINATION's inbox with one atomicAdd on that
//     rank's cursor (a remote atomic over NVLink for a peer: one per 1024 list entries and destination; same-address
//     atomics complete at a few ns each, so their number, not their latency, is what a scan pays for) and is then
//     written there with coalesced stores.  A voxel with hits and misses has two list entries: the hit entry handles both.
//     Clears the local staging it consumes;
// (3) the last CTA (completion ticket) publishes the counts to every destination's mailbox, makes everything visible
//     system-wide and raises this source's epoch flag in every arena, then rearms the local staging counters for the
//     owner-side ingest.
__global__ void __launch_bounds__(kPushThreads) k_shard_push(ShardPeers X, MapParams P, DeviceBuffers D, FrameParams F, int par,
                                                             int *sent /*[kMaxWorld + 1] is the completion ticket*/, uint32_t epoch) {
  __shared__ int s_cnt[kMaxWorld];
  __shared__ int s_base[kMaxWorld];
  __shared__ int s_goff[kMaxWorld + 1];
  __shared__ ShardRecord s_stage[kPushStage];
  __shared__ int s_last;
  FrameCounters *fc = D.fc[F.parity];
  const int world = X.world;
  // F is the owner-side view of the frame (stage_only off, inline_resolve on): the voxels that stay resolve their subboxes here
  // (1)
  const int n_hit = fc->n_hit;
  if (n_hit > X.hit_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) fc->error = kErrCapacity;
  } else {
    const size_t region = ((size_t)par * world + X.rank) * X.hit_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hit; i += gridDim.x * blockDim.x) {
      const int key = D.hit_key[i];
      const uint32_t st = D.hit_t[i];
      for (int d = 0; d < world; d++) {
        X.a[d].gather_key[region + i] = key;
        X.a[d].gather_stamp[region + i] = st;
      }
    }
  }
  // (2)
  const int n = min(fc->n_touched_remote, P.max_touched);
  const int dxy = P.lvg_dim_xy;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    if (threadIdx.x < world) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = base + threadIdx.x;
    int dest = -1, nrec = 0, head = kLvgEmpty, mc = 0, lv = 0, rank_in_cta = 0;
    int c[3] = {0, 0, 0};
    if (i < n) {
      const uint32_t e = D.touched_remote[i];
      lv = (int)(e & ~kTouchedHitTag);
      const int2 st = D.lvg[lv];
      if ((e & kTouchedHitTag) || st.x == kLvgEmpty) {  // the entry that speaks for the voxel
        head = st.x;
        mc = st.y;
        const int lz = (int)fast_div((uint32_t)lv, P.dxy2_mul, P.dxy2_shift);
        const int rem = lv - lz * dxy * dxy;
        const int ly = (int)fast_div((uint32_t)rem, P.dxy_mul, P.dxy_shift);
        c[0] = rem - ly * dxy + F.lvg_base[0];
        c[1] = ly + F.lvg_base[1];
        c[2] = lz + F.lvg_base[2];
        int g[3];
        dest = owner_of(P, c, world, g);
        for (int h = head; h != kLvgEmpty; h = D.hit_next[h]) nrec++;
        if (mc > 0) nrec++;
        rank_in_cta = atomicAdd(&s_cnt[dest], nrec);
      }
    }
    __syncthreads();
    if (threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(X.a[threadIdx.x].cursor + par, s_cnt[threadIdx.x]);
    // records go to shared memory first, grouped by destination, and leave as contiguous runs of coalesced 8-byte
    // stores (scattered 24-byte records written thread by thread cost 60 us per scan over NVLink)
    if (threadIdx.x == 32) {
      int acc = 0;
      for (int d = 0; d < world; d++) {
        s_goff[d] = acc;
        acc += s_cnt[d];
      }
      s_goff[world] = acc;
    }
    __syncthreads();
    const int chunk_total = s_goff[world];
    const bool staged = chunk_total <= kPushStage;   // a chunk with unusually long hit lists writes straight through
    if (dest >= 0) {
      const int first = s_base[dest] + rank_in_cta;
      if (first + nrec > X.rec_cap) {
        fc->error = kErrCapacity;
      } else {
        ShardRecord *o = staged ? s_stage + s_goff[dest] + rank_in_cta : X.a[dest].inbox + (size_t)par * X.rec_cap + first;
        for (int h = head; h != kLvgEmpty; h = D.hit_next[h]) {
          ShardRecord r;
          r.c[0] = c[0];
          r.c[1] = c[1];
          r.c[2] = c[2];
          r.key = D.hit_key[h];
          r.p = D.hit_p[h];
          r.count = (int)D.hit_t[h];  // every key is cast by exactly one rank, so its stamp travels with it
          *o++ = r;
        }
        if (mc > 0) {
          ShardRecord r;
          r.c[0] = c[0];
          r.c[1] = c[1];
          r.c[2] = c[2];
          r.key = -1;
          r.p = 0.f;
          r.count = mc;
          *o++ = r;
        }
      }
      // staging consumed.  Only the entry that owns the voxel clears it: the other entry of a voxel with hits and
      // misses may sit anywhere in the list, and clearing from there could empty the voxel before its owner reads it
      D.lvg[lv] = make_int2(kLvgEmpty, 0);
    }
    __syncthreads();
    if (staged) {
      const int2 *src = reinterpret_cast<const int2 *>(s_stage);
      for (int d = 0; d < world; d++) {
        const int cnt = s_cnt[d];
        if (cnt == 0 || s_base[d] + cnt > X.rec_cap) continue;
        int2 *dst = reinterpret_cast<int2 *>(X.a[d].inbox + (size_t)par * X.rec_cap + s_base[d]);
        const int2 *sd = src + 3 * s_goff[d];
        for (int q = threadIdx.x; q < 3 * cnt; q += blockDim.x) dst[q] = sd[q];
      }
    }
    __syncthreads();
  }
  // (3) ONE system-scope fence per CTA, after the CTA barrier: fences are cumulative, so it orders the (remote) stores of
  // every thread of the CTA (a fence.sys in every thread serialises inside the SM: ~40 us per scan); then the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(&sent[kMaxWorld + 1], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int d = threadIdx.x;
  if (d == X.rank) {   // the entries this rank staged for itself, for its owner-side kernel (the thread that raises this
                       // rank's own flag writes them, ahead of its fence)
    sent[kMaxWorld + 2] = n_hit;
    sent[kMaxWorld + 3] = __ldcg(&fc->n_touched);
  }
  if (d < world) {
    const bool failed = __ldcg(&fc->error) != 0;
    X.a[d].mbox[par * kMaxWorld + X.rank] = make_int2(failed ? -1 : n_hit, failed ? -1 : 0);
    // (the release orders this thread's stores above and, through the tickets behind the CTAs' fences, everyone else's)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(X.a[d].flags + X.rank), "r"(epoch) : "memory");
  }
}

// ---- the scan itself: every rank copies 1/world of it from the host, the slices travel over NVLink ------------------
// (a scan of 262 k points is 6.3 MB: copied whole by every rank it costs 125-210 us of PCIe time per scan and rank)
// source: own slice [first, first + n) of points[par] (already in this rank's arena) -> the same place in every peer's
// arena with 16-byte stores; the last CTA raises this rank's point flag everywhere.  state[0] = ticket
__global__ void __launch_bounds__(256) k_shard_scatter_points(ShardPeers X, int par, int first, int n, uint32_t epoch, int *state) {
  __shared__ int s_last;
  const size_t base = (size_t)par * 3 * X.max_points;
  // 3 doubles per point: the slice is a run of 8-byte words [3 * first, 3 * (first + n))
  const double *src = X.a[X.rank].points + base + 3 * (size_t)first;
  const size_t words = 3 * (size_t)n;
  for (int r = 0; r < X.world; r++) {
    if (r == X.rank) continue;
    double *dst = X.a[r].points + base + 3 * (size_t)first;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(&state[0], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) state[0] = 0;
  if (threadIdx.x < X.world)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(X.a[threadIdx.x].pflags + X.rank), "r"(epoch) : "memory");
}
// every rank: all slices of the scan have arrived (bounded spin; *err = kErrPeer otherwise, which fails the scan)
__global__ void k_shard_wait_points(ShardPeers X, uint32_t epoch, unsigned long long timeout_ns, FrameCounters *fc) {
  if ((int)threadIdx.x >= X.world) return;
  const uint32_t *flag = X.a[X.rank].pflags + threadIdx.x;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int)(v - epoch) >= 0) return;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      fc->error = kErrPeer;
      return;
    }
    __nanosleep(100);
  }
}

// ---- owner side ---------------------------------------------------------------------------------------------------
// wait until every source has raised its flag for this epoch, total the counts and decide whether the scan crosses a
// libstdc++ rehash (same decision on every rank: they all see the same counts).  Every CTA of the first owner-side
// kernel runs this (the flags are read-only here); CTA 0 records the result for the later kernels and the host.
__device__ __forceinline__ bool shard_wait(const ShardPeers &X, DeviceBuffers &D, const FrameParams &F, int par, uint32_t epoch,
                                           ShardState *st, int *skip, unsigned long long timeout_ns, ShardState *s_st,
                                           int P_max_touched, const int *local_counts) {
  __shared__ int s_fail;
  const int s = threadIdx.x;
  if (s == 0) s_fail = 0;
  __syncthreads();
  unsigned long long t0, t1 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (s < X.world) {
    const uint32_t *flag = X.a[X.rank].flags + s;
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if ((int)(v - epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        s_fail = 1;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (s == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    FrameCounters *fc = D.fc[F.parity];
    int n_total = 0, n_rec = 0, err = s_fail ? kErrPeer : 0;
    for (int r = 0; r < X.world; r++) {
      int2 m = s_fail ? make_int2(0, 0) : __ldcg(&X.a[X.rank].mbox[par * kMaxWorld + r]);
      if (m.x < 0 || m.y < 0) {
        err = kErrPeer;
        m = make_int2(0, 0);
      }
      s_st->cnt_hits[r] = m.x;
      n_total += m.x;
    }
    // every source has reserved and written all its records: the cursor is the number of records to ingest
    n_rec = __ldcg(X.a[X.rank].cursor + par);
    if (!err && n_rec > X.rec_cap) err = kErrCapacity;  // a source ran past rec_cap (it flagged its own error too)
    if (err) n_rec = 0;
    const int local_err = __ldcg(&fc->error);
    if (local_err) err = local_err;
    s_st->n_total = n_total;
    s_st->n_rec_total = n_rec;
    // (as the push kernel left them: the ingest part of this kernel moves fc->n_hit / n_touched while later CTAs still wait)
    s_st->hit_base = __ldcg(local_counts);
    s_st->touched_base = min(__ldcg(local_counts + 1), P_max_touched);
    s_st->rehash = (err == 0 && (uint32_t)n_total > F.bucket_count) ? 1 : 0;
    s_st->error = err;
    s_st->wait_ns = t1 - t0;
    if (blockIdx.x == 0) {
      *st = *s_st;
      if (err) fc->error = err;
      *skip = (err != 0 || s_st->rehash) ? 1 : 0;
    }
  }
  __syncthreads();
  return s_st->error == 0 && s_st->rehash == 0;
}

// rehash scans: the hit entries this rank staged itself get their virtual positions and buckets like the ingested ones
__global__ void k_shard_restamp(MapParams P, DeviceBuffers D, int n_local, const uint32_t *key_stamp, uint32_t B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const int key = D.hit_key[i];
  D.hit_t[i] = key_stamp[key];
  D.hit_bucket[i] = cell_bucket(P, key, B, 0);
}
// global ordering info per key: key_stamp[key] = first-insert stamp (or virtual position on a rehash frame)
__global__ void k_shard_scatter_stamps(const int *keys, const uint32_t *stamps, int n, uint32_t *key_stamp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key_stamp[keys[i]] = stamps[i];
}
// bucket activation stamps of a no-rehash scan from the gathered (key, stamp) lists of all sources
__device__ __forceinline__ void act_body(const ShardPeers &X, const MapParams &P, const FrameParams &F, int par, const int *cnt_hits,
                                         uint32_t *act) {
  const ShardArena &A = X.a[X.rank];
  const uint32_t B = F.bucket_count;
  for (int src = 0; src < X.world; src++) {
    const int n = cnt_hits[src];
    const size_t region = ((size_t)par * X.world + src) * X.hit_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      atomicMin(&act[hit_bucket_fast(P, __ldcg(A.gather_key + region + i), B, F.bucket_c64)], __ldcg(A.gather_stamp + region + i));
  }
}

// owner side: received records -> hit arrays + voxel-grid staging (the role k_column's staging plays on one GPU).
// No list-building atomics: record i of the inbox owns slot i of the frame's touched list (an invalid marker when the
// voxel was staged before) and, while the inbox fits the hit list, slot i of the hit arrays; k_fuse skips the markers.
// The first toucher of a subbox resolves / allocates it on the spot (F.inline_resolve, as in k_frame).
constexpr uint32_t kTouchedNone = 0xffffffffu;
__device__ __forceinline__ void ingest_body(const ShardPeers &X, const MapParams &P, DeviceBuffers &D, const FrameParams &F, int par,
                                            int n, int hb, int tb, const uint32_t *key_stamp, int *s_warp) {
  FrameCounters *fc = D.fc[F.parity];
  if (tb + n > P.max_touched) {   // the voxels this rank staged itself come first
    if (blockIdx.x == 0 && threadIdx.x == 0) fc->error = kErrCapacity;
    return;
  }
  const bool dense_hits = hb + n <= P.max_hits;   // uniform over the grid
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    fc->n_touched = tb + n;
    if (dense_hits) fc->n_hit = hb + n;
  }
  const int2 *rec = reinterpret_cast<const int2 *>(X.a[X.rank].inbox + (size_t)par * X.rec_cap);
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    const bool valid = i < n;
    ShardRecord r;
    r.key = -1;
    int lv = -1;
    CellRef cr;
    if (valid) {
      const int2 a = __ldcg(rec + 3 * (size_t)i), b = __ldcg(rec + 3 * (size_t)i + 1), c = __ldcg(rec + 3 * (size_t)i + 2);
      r.c[0] = a.x;
      r.c[1] = a.y;
      r.c[2] = b.x;
      r.key = b.y;
      r.p = __int_as_float(c.x);
      r.count = c.y;
      for (int k = 0; k < 3; k++) {
        cr.c[k] = r.c[k];
        cr.g[k] = fast_floor_div(r.c[k], P.n, P.n_mul, P.n_shift);
      }
      lv = lvg_index(P, F, cr);
      if (lv < 0) fc->error = kErrInternal;
    }
    const bool is_hit = valid && lv >= 0 && r.key >= 0;
    const bool is_miss = valid && lv >= 0 && r.key < 0;
    int idx = hb + i;
    if (!dense_hits) idx = cta_reserve(&fc->n_hit, is_hit ? 1 : 0, s_warp);   // CTA-uniform branch
    uint32_t entry = kTouchedNone;
    if (is_hit) {
      if (idx >= P.max_hits) {
        fc->error = kErrCapacity;
      } else {
        D.hit_key[idx] = r.key;
        D.hit_p[idx] = r.p;
        D.hit_t[idx] = key_stamp ? key_stamp[r.key] : (uint32_t)r.count;  // rehash scans: virtual position from the global order
        D.hit_bucket[idx] = hit_bucket_fast(P, r.key, F.bucket_count, F.bucket_c64);
        const int old = atomicExch(&D.lvg[lv].x, idx);
        D.hit_next[idx] = old;
        if (old == kLvgEmpty) entry = (uint32_t)lv | kTouchedHitTag;
      }
    } else if (is_miss) {
      if (atomicAdd(&D.lvg[lv].y, r.count) == 0) entry = (uint32_t)lv;
    }
    if (valid) {
      D.touched[tb + i] = entry;
      if (lv >= 0) touch_subbox(P, F, D, fc, cr.g);
    }
  }
}

// the owner-side kernel of a no-rehash scan: every CTA waits for the sources (the flags are read-only here), then the
// grid derives the bucket activations from the gathered keys and stages the received records; the two parts touch
// disjoint data, so they need no barrier between them.  A scan that needs the rehash path (or failed) returns at once,
// with `skip` raised for k_fuse.
__global__ void __launch_bounds__(256) k_shard_act_ingest(ShardPeers X, MapParams P, DeviceBuffers D, FrameParams F, int par, uint32_t epoch,
                                                          ShardState *st, int *skip, unsigned long long timeout_ns, uint32_t *act) {
  __shared__ ShardState s_st;
  __shared__ int s_warp[33];
  if (!shard_wait(X, D, F, par, epoch, st, skip, timeout_ns, &s_st, P.max_touched, skip + 2)) return;
  act_body(X, P, F, par, s_st.cnt_hits, act);
  ingest_body(X, P, D, F, par, s_st.n_rec_total, s_st.hit_base, s_st.touched_base, nullptr, s_warp);
}

// rehash scans (host-driven, mlmap_capi.cu shard_rehash_path): the records are staged with the virtual positions the
// re-sequencing assigned
__global__ void __launch_bounds__(256) k_shard_ingest(ShardPeers X, MapParams P, DeviceBuffers D, FrameParams F, int par,
                                                      const ShardState *st, const uint32_t *key_stamp) {
  __shared__ int s_warp[33];
  ingest_body(X, P, D, F, par, st->n_rec_total, st->hit_base, st->touched_base, key_stamp, s_warp);
}

// ---- replicated map (query stream): after a frame the owner ships the subbox blocks the frame
// touched ("dirty" blocks); replicas overwrite / create them.  Record = 16-byte header {g[3], cells} followed
// by log_odds[cells] f32, occupancy[cells], inflate[cells] (padded to 16 bytes).
__host__ __device__ inline size_t dirty_record_bytes(int cells) { return (16 + (size_t)cells * 6 + 15) & ~(size_t)15; }

// record i of the frame's dirty list -> rec
__device__ __forceinline__ void dirty_pack_record(const MapParams &P, const DeviceBuffers &D, const FrameParams &F, int i, unsigned char *rec) {
  const int ls = D.touched_sub[i];
  const int block = D.lsg_block[ls];
  int *hdr = reinterpret_cast<int *>(rec);
  if (threadIdx.x == 0) {
    int lx = ls % P.lsg_dim_xy, ly = (ls / P.lsg_dim_xy) % P.lsg_dim_xy, lz = ls / (P.lsg_dim_xy * P.lsg_dim_xy);
    hdr[0] = lx + F.lsg_base[0];
    hdr[1] = ly + F.lsg_base[1];
    hdr[2] = lz + F.lsg_base[2];
    hdr[3] = block >= 0 ? P.cells : 0;  // 0: nothing to ship (collapsed / unusable)
  }
  if (block < 0) return;
  float *lo = reinterpret_cast<float *>(rec + 16);
  char *occ = reinterpret_cast<char *>(rec + 16 + (size_t)P.cells * 4);
  char *inf = occ + P.cells;
  const size_t src = (size_t)block * P.cell_stride;
  for (int c = threadIdx.x; c < P.cells; c += blockDim.x) {
    lo[c] = D.pool_lo[src + c];
    occ[c] = D.pool_occ[src + c];
    inf[c] = D.pool_inf[src + c];
  }
}

__global__ void __launch_bounds__(256) k_dirty_export(MapParams P, DeviceBuffers D, FrameParams F, int n, unsigned char *out) {
  const int i = blockIdx.x;
  if (i >= n) return;
  dirty_pack_record(P, D, F, i, out + (size_t)i * dirty_record_bytes(P.cells));
}

// rec -> the replica's map: the subbox is found or created (hash insert + free-stack pop by thread 0), then overwritten.
// Called by a whole CTA; counters[0] += new blocks, counters[1] = error
__device__ __forceinline__ void dirty_apply_record(const MapParams &P, DeviceBuffers &D, const unsigned char *rec, int *counters, int *s_block) {
  const int *hdr = reinterpret_cast<const int *>(rec);
  if (hdr[3] == 0) return;   // (uniform over the CTA)
  if (threadIdx.x == 0) {
    int g[3] = {hdr[0], hdr[1], hdr[2]};
    uint64_t key;
    int block = kBlockUnusable;
    if (pack_glb(g, key)) {
      uint32_t slot = ht_hash(key) & P.ht_mask;
      for (uint32_t probe = 0; probe <= P.ht_mask; probe++) {
        uint64_t k = D.ht_key[slot];
        if (k == kEmptyKey) {
          unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(&D.ht_key[slot]),
                                             (unsigned long long)kEmptyKey, (unsigned long long)key);
          if (old == kEmptyKey) {
            int top = atomicSub(D.free_top, 1) - 1;
            if (top < 0) {
              atomicAdd(D.free_top, 1);
              D.ht_val[slot] = kBlockUnusable;
              counters[1] = kErrPool;
            } else {
              block = D.free_stack[top];
              D.ht_val[slot] = block;
              atomicAdd(&counters[0], 1);
            }
            break;
          }
          k = (uint64_t)old;
        }
        if (k == key) {
          block = D.ht_val[slot];
          break;
        }
        slot = (slot + 1) & P.ht_mask;
      }
    }
    *s_block = block;
  }
  __syncthreads();
  const int block = *s_block;
  if (block < 0) return;
  const float *lo = reinterpret_cast<const float *>(rec + 16);
  const char *occ = reinterpret_cast<const char *>(rec + 16 + (size_t)P.cells * 4);
  const char *inf = occ + P.cells;
  const size_t dst = (size_t)block * P.cell_stride;
  for (int c = threadIdx.x; c < P.cells; c += blockDim.x) {
    D.pool_lo[dst + c] = lo[c];
    D.pool_occ[dst + c] = occ[c];
    D.pool_inf[dst + c] = inf[c];
  }
}

__global__ void __launch_bounds__(256) k_dirty_import(MapParams P, DeviceBuffers D, int n, const unsigned char *in, int *counters) {
  __shared__ int s_block;
  const int i = blockIdx.x;
  if (i >= n) return;
  dirty_apply_record(P, D, in + (size_t)i * dirty_record_bytes(P.cells), counters, &s_block);
}

// ---- the same replication with the records stored straight into the replicas' arenas over NVLink peer memory --------
// Every rank of a replicated map owns one arena of the same layout: inbox[2][cap] (dirty records of a frame, by frame
// parity), count[2], flags[2] (the frame number whose records sit in that half; written by the source with a system-
// scope release), ack[world] (in the SOURCE's arena: the last frame each replica has applied, written by the replica).
// The source may run at most two frames ahead of the slowest replica: before it overwrites a half it waits for the acks
// of the frame that used it.
struct ReplicaArena {
  uint32_t *flags;
  int *count;
  uint32_t *ack;
  unsigned char *inbox;
};
struct ReplicaPeers {
  int rank, world, src, cap_blocks;
  unsigned long long half_bytes;   // bytes of one inbox half
  ReplicaArena a[kMaxWorld];
};

__device__ __forceinline__ bool spin_until(const uint32_t *word, uint32_t target, unsigned long long timeout_ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
    if ((int)(v - target) >= 0) return true;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) return false;
    __nanosleep(100);
  }
}

// source: one CTA per dirty block packs it into every replica's inbox; the last CTA publishes the count and the flag.
// state[0] = completion ticket, state[1] = error
__global__ void __launch_bounds__(256) k_replica_push(ReplicaPeers X, MapParams P, DeviceBuffers D, FrameParams F, int n, uint32_t epoch,
                                                      int *state, unsigned long long timeout_ns) {
  __shared__ int s_fail, s_last;
  const int par = (int)(epoch & 1);
  if (threadIdx.x == 0) s_fail = 0;
  __syncthreads();
  if (epoch > 2 && threadIdx.x < X.world && threadIdx.x != X.src)
    if (!spin_until(X.a[X.src].ack + threadIdx.x, epoch - 2, timeout_ns)) s_fail = 1;
  __syncthreads();
  const bool failed = s_fail != 0;
  if (!failed && (int)blockIdx.x < n) {
    const size_t off = (size_t)par * X.half_bytes + (size_t)blockIdx.x * dirty_record_bytes(P.cells);
    for (int r = 0; r < X.world; r++)
      if (r != X.src) dirty_pack_record(P, D, F, (int)blockIdx.x, X.a[r].inbox + off);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (failed) state[1] = kErrPeer;
    __threadfence_system();
    s_last = atomicAdd(&state[0], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) state[0] = 0;
  const int d = threadIdx.x;
  if (d < X.world && d != X.src) {
    X.a[d].count[par] = __ldcg(&state[1]) ? -1 : n;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(X.a[d].flags + par), "r"(epoch) : "memory");
  }
}

// replica: wait for the frame's flag, apply its records, acknowledge.  counters: [0] new blocks, [1] error, [2] ticket, [3] records
__global__ void __launch_bounds__(256) k_replica_pull(ReplicaPeers X, MapParams P, DeviceBuffers D, uint32_t epoch, int *counters,
                                                      unsigned long long timeout_ns) {
  __shared__ int s_n, s_last, s_block;
  const int par = (int)(epoch & 1);
  const ReplicaArena &A = X.a[X.rank];
  if (threadIdx.x == 0) {
    int n = -1;
    if (spin_until(A.flags + par, epoch, timeout_ns)) n = __ldcg(A.count + par);
    if (n < 0 || n > X.cap_blocks) {
      counters[1] = kErrPeer;
      n = 0;
    }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  const unsigned char *in = A.inbox + (size_t)par * X.half_bytes;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    dirty_apply_record(P, D, in + (size_t)i * dirty_record_bytes(P.cells), counters, &s_block);
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&counters[2], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  counters[2] = 0;
  counters[3] = n;
  // the half may be overwritten once every CTA has read it: tell the source
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(X.a[X.src].ack + X.rank), "r"(epoch) : "memory");
}


// ---- checkpoint / restore of the submap pool (the reference has no persistence) -------------------
// One fixed-size record per subbox: 32-byte header {g[3], flags, element 0 of a collapsed subbox} followed by
// log_odds[cells] f32, occupancy[cells], inflate_occupancy[cells] and, in exploration mode, the frontier bitmask.
struct CkptRecHeader {
  int g[3];
  int flags;        // 0: full block follows, 1: collapsed subbox (only col_* are meaningful, payload is zero)
  float col_lo;
  char col_occ, col_inf;
  char pad[10];
};
static_assert(sizeof(CkptRecHeader) == 32, "checkpoint record header is 32 bytes");
__host__ __device__ inline size_t ckpt_record_bytes(int cells, int front_words, int explore) {
  return (sizeof(CkptRecHeader) + (size_t)cells * 6 + (explore ? (size_t)front_words * 4 : 0) + 15) & ~(size_t)15;
}
// blocks[i] as produced by k_export_list: pool block, or -16 - hash slot for a collapsed subbox
__global__ void __launch_bounds__(256) k_ckpt_pack(MapParams P, DeviceBuffers D, const int *glb3, const int *blocks, int first,
                                                   int n, unsigned char *out) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const size_t rb = ckpt_record_bytes(P.cells, P.front_words, P.explore);
  unsigned char *rec = out + (size_t)i * rb;
  const int blk = blocks[first + i];
  for (size_t b = threadIdx.x * 16; b < rb; b += blockDim.x * 16) *reinterpret_cast<uint4 *>(rec + b) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  CkptRecHeader *hd = reinterpret_cast<CkptRecHeader *>(rec);
  if (threadIdx.x == 0) {
    hd->g[0] = glb3[3 * (first + i)];
    hd->g[1] = glb3[3 * (first + i) + 1];
    hd->g[2] = glb3[3 * (first + i) + 2];
    hd->flags = blk <= -16 ? 1 : 0;
    if (blk <= -16) {
      const int slot = -(blk + 16);
      hd->col_lo = D.col_lo[slot];
      hd->col_occ = D.col_occ[slot];
      hd->col_inf = D.col_inf[slot];
    }
  }
  if (blk < 0) return;
  float *lo = reinterpret_cast<float *>(rec + sizeof(CkptRecHeader));
  char *occ = reinterpret_cast<char *>(lo + P.cells);
  char *inf = occ + P.cells;
  const size_t src = (size_t)blk * P.cell_stride;
  for (int c = threadIdx.x; c < P.cells; c += blockDim.x) {
    lo[c] = D.pool_lo[src + c];
    occ[c] = D.pool_occ[src + c];
    inf[c] = D.pool_inf[src + c];
  }
  if (P.explore) {
    // the frontier words sit behind the three cell arrays, 4-byte aligned because cells*6 is even... keep it exact:
    unsigned char *fw = reinterpret_cast<unsigned char *>(inf + P.cells);
    for (int w = threadIdx.x; w < P.front_words; w += blockDim.x) {
      const uint32_t v = D.pool_front[(size_t)blk * P.front_words + w];
      fw[4 * w] = (unsigned char)v;
      fw[4 * w + 1] = (unsigned char)(v >> 8);
      fw[4 * w + 2] = (unsigned char)(v >> 16);
      fw[4 * w + 3] = (unsigned char)(v >> 24);
    }
  }
}
// inverse: insert the subbox into the (freshly reset) hash table, take a block from the free stack and fill it
__global__ void __launch_bounds__(256) k_ckpt_unpack(MapParams P, DeviceBuffers D, int n, const unsigned char *in, int *status) {
  __shared__ int s_block;
  const int i = blockIdx.x;
  if (i >= n) return;
  const size_t rb = ckpt_record_bytes(P.cells, P.front_words, P.explore);
  const unsigned char *rec = in + (size_t)i * rb;
  const CkptRecHeader *hd = reinterpret_cast<const CkptRecHeader *>(rec);
  if (threadIdx.x == 0) {
    int block = kBlockUnusable;
    int g[3] = {hd->g[0], hd->g[1], hd->g[2]};
    uint64_t key;
    if (!pack_glb(g, key)) {
      *status = kErrRange;
    } else {
      uint32_t slot = ht_hash(key) & P.ht_mask;
      bool done = false;
      for (uint32_t probe = 0; probe <= P.ht_mask && !done; probe++) {
        unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long *>(&D.ht_key[slot]),
                                           (unsigned long long)kEmptyKey, (unsigned long long)key);
        if (old == kEmptyKey) {
          if (hd->flags & 1) {
            if (P.explore) {
              D.col_lo[slot] = hd->col_lo;
              D.col_occ[slot] = hd->col_occ;
              D.col_inf[slot] = hd->col_inf;
              D.ht_val[slot] = kBlockCollapsed;
              block = kBlockCollapsed;
            } else {
              *status = kErrInternal;  // a collapsed subbox cannot exist without the exploration mode
            }
          } else {
            const int top = atomicSub(D.free_top, 1) - 1;
            if (top < 0) {
              atomicAdd(D.free_top, 1);
              D.ht_val[slot] = kBlockUnusable;
              *status = kErrPool;
            } else {
              block = D.free_stack[top];
              D.ht_val[slot] = block;
            }
          }
          done = true;
        } else if (old == key) {
          *status = kErrInternal;  // duplicate subbox in the checkpoint
          done = true;
        } else {
          slot = (slot + 1) & P.ht_mask;
        }
      }
      if (!done) *status = kErrPool;
    }
    s_block = block;
  }
  __syncthreads();
  const int block = s_block;
  if (block < 0) return;
  const float *lo = reinterpret_cast<const float *>(rec + sizeof(CkptRecHeader));
  const char *occ = reinterpret_cast<const char *>(lo + P.cells);
  const char *inf = occ + P.cells;
  const size_t dst = (size_t)block * P.cell_stride;
  for (int c = threadIdx.x; c < P.cells; c += blockDim.x) {
    D.pool_lo[dst + c] = lo[c];
    D.pool_occ[dst + c] = occ[c];
    D.pool_inf[dst + c] = inf[c];
  }
  if (P.explore) {
    const unsigned char *fw = reinterpret_cast<const unsigned char *>(inf + P.cells);
    for (int w = threadIdx.x; w < P.front_words; w += blockDim.x)
      D.pool_front[(size_t)block * P.front_words + w] =
          (uint32_t)fw[4 * w] | ((uint32_t)fw[4 * w + 1] << 8) | ((uint32_t)fw[4 * w + 2] << 16) | ((uint32_t)fw[4 * w + 3] << 24);
  }
}

}  // namespace mlm
