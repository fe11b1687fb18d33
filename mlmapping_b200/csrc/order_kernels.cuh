// Emulation of libstdc++'s unordered_map iteration order for frames that cross a rehash
// (/usr/include/c++/13/bits/hashtable.h:2009-2033 insert-at-bucket-begin,
// :2584-2620 _M_rehash_aux re-inserts the nodes in their current iteration order).
//
// The reference consumes hit_idx_odds_hashmap in iteration order (src/map_local.cpp:147) and
// the clamped hit update is order dependent, so bit-exact log-odds need that order.  With B
// buckets and no rehash during the frame the order is: descending (activation stamp of the key's
// bucket, first-insert stamp of the key) — handled inside k_column/k_fuse with two atomicMin
// stamps.  When the number of distinct keys exceeds B, libstdc++ rehashes mid-frame: the current
// iteration order becomes the virtual insertion sequence for the next bucket count.  This file
// implements that chain with data-parallel stages: stamp -> sort -> re-sequence.
#pragma once
#include "frame_kernels.cuh"

namespace mlm {

constexpr int kSortChunk = 4096;  // elements sorted per CTA in shared memory
constexpr int kSortThreads = 512;

// global compare-exchange step of a bitonic network (ascending overall)
__global__ void __launch_bounds__(256) k_bitonic_global(uint64_t *keys, int n_pad, int j, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (n_pad >> 1)) return;
  int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
  int b = a | j;
  uint64_t ka = keys[a], kb = keys[b];
  bool up = (a & k) == 0;
  if ((ka > kb) == up) {
    keys[a] = kb;
    keys[b] = ka;
  }
}

// all steps with j < kSortChunk for k in [k_lo, k_hi] on one chunk per CTA, in shared memory.
// k_lo == 2 sorts the chunk from scratch; k_lo == k_hi > kSortChunk finishes one merge level.
__global__ void __launch_bounds__(kSortThreads) k_bitonic_local(uint64_t *keys, int n_pad, int k_lo, int k_hi) {
  __shared__ uint64_t s[kSortChunk];
  const int chunk = min(kSortChunk, n_pad);
  const int base = blockIdx.x * chunk;
  for (int i = threadIdx.x; i < chunk; i += blockDim.x) s[i] = keys[base + i];
  __syncthreads();
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = min(k >> 1, chunk >> 1); j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < (chunk >> 1); i += blockDim.x) {
        int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        int b = a | j;
        uint64_t ka = s[a], kb = s[b];
        bool up = ((base + a) & k) == 0;
        if ((ka > kb) == up) {
          s[a] = kb;
          s[b] = ka;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < chunk; i += blockDim.x) keys[base + i] = s[i];
}

// ---- stage kernels ------------------------------------------------------------------------------------
// bucket of an awareness cell index in the emulated container:
//   kind 0: hit_idx_odds_hashmap, key Vec3I(rho,phi,z), VectorHasher (include/map_awareness.h:31-41,56)
//   kind 1: miss_idx_set, key size_t mapIdx, std::hash<size_t> = identity (include/map_awareness.h:57)
__device__ __forceinline__ uint32_t cell_bucket(const MapParams &P, int key, uint32_t B, int kind) {
  if (kind == 1) return (uint32_t)((uint64_t)(uint32_t)key % (uint64_t)B);
  int zk = key / (P.nRho * P.nPhi), rem = key - zk * (P.nRho * P.nPhi);
  int pk = rem / P.nRho, rk = rem - pk * P.nRho;
  return libstdcxx_bucket(vector_hash3(rk, pk, zk), B);
}
// hit-map bucket at the frame's bucket count, 32-bit arithmetic only (c64 = 2^64 mod B from the host)
__device__ __forceinline__ uint32_t hit_bucket_fast(const MapParams &P, int key, uint32_t B, uint32_t c64) {
  const int zk = key / (P.nRho * P.nPhi), rem = key - zk * (P.nRho * P.nPhi);
  const int pk = rem / P.nRho, rk = rem - pk * P.nRho;
  return libstdcxx_bucket_fast(vector_hash3(rk, pk, zk), B, c64);
}
struct OrderArrays {
  const int *key;      // awareness cell index of every element
  uint32_t *stamp;     // in: first-insert stamps; out (slow path): virtual positions
  uint32_t *bucket;    // out: bucket at the final bucket count
  int kind;
};

// keys[i] = (first-insert stamp, hit index): ascending sort = real insertion sequence
__global__ void k_order_seed(OrderArrays O, uint64_t *keys, int n, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    keys[i] = ((uint64_t)O.stamp[i] << 32) | (uint32_t)i;
  else if (i < n_pad)
    keys[i] = ~0ull;
}
__global__ void k_order_take_seq(const uint64_t *keys, int *seq, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seq[i] = (int)(uint32_t)keys[i];
}
__global__ void k_fill_u32(uint32_t *p, uint32_t v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// act[bucket] = earliest virtual position among the first m keys of the sequence
__global__ void k_stage_act(MapParams P, OrderArrays O, uint32_t *act, const int *seq, int m, uint32_t B) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) atomicMin(&act[cell_bucket(P, O.key[seq[j]], B, O.kind)], (uint32_t)j);
}
// sort key: descending (act, position)  ==  ascending 63-bit complement; padding sorts last
__global__ void k_stage_keys(MapParams P, OrderArrays O, const uint32_t *act, const int *seq, uint64_t *keys, int m,
                             int m_pad, uint32_t B) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) {
    uint64_t comp = ((uint64_t)act[cell_bucket(P, O.key[seq[j]], B, O.kind)] << 32) | (uint32_t)j;
    keys[j] = (~comp) & 0x7fffffffffffffffull;
  } else if (j < m_pad) {
    keys[j] = ~0ull;
  }
}
__global__ void k_stage_apply(const uint64_t *keys, const int *seq_old, int *seq_new, int m) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) {
    uint32_t old_pos = (uint32_t)((~keys[j]) & 0xffffffffull);
    seq_new[j] = seq_old[old_pos];
  }
}
__global__ void k_copy_i32(int *dst, const int *src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
// final: stamps become virtual positions; bucket activation over the whole sequence
__global__ void k_order_final(MapParams P, OrderArrays O, uint32_t *act, const int *seq, int n, uint32_t B) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    int h = seq[j];
    O.stamp[h] = (uint32_t)j;
    const uint32_t b = cell_bucket(P, O.key[h], B, O.kind);
    O.bucket[h] = b;
    atomicMin(&act[b], (uint32_t)j);
  }
}

// ---- export of the hit map in iteration order (parity/debug) -------------------------------------------
// keys_t : ascending (stamp, hit index) from k_order_seed + sort.  keys_o : ascending complement of
// (bucket activation, stamp) = the iteration order.  The stamp is unique per key, so the hit index of
// an ordered entry is recovered by binary search of its stamp in keys_t.
__global__ void k_export_hit_keys(MapParams P, DeviceBuffers D, const uint32_t *act, uint64_t *keys_o, int n, int n_pad,
                                  uint32_t B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t comp = ((uint64_t)act[cell_bucket(P, D.hit_key[i], B, 0)] << 32) | D.hit_t[i];
    keys_o[i] = (~comp) & 0x7fffffffffffffffull;
  } else if (i < n_pad) {
    keys_o[i] = ~0ull;
  }
}
__global__ void k_export_hit_gather(MapParams P, DeviceBuffers D, const uint64_t *keys_t, const uint64_t *keys_o,
                                    int n, int *out_key3, float *out_p) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t stamp = (uint32_t)((~keys_o[j]) & 0xffffffffull);
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((uint32_t)(keys_t[mid] >> 32) < stamp) lo = mid + 1; else hi = mid;
  }
  int h = (int)(uint32_t)keys_t[lo];
  int key = D.hit_key[h];
  int zk = key / (P.nRho * P.nPhi), rem = key - zk * (P.nRho * P.nPhi);
  int pk = rem / P.nRho, rk = rem - pk * P.nRho;
  out_key3[3 * j] = rk;
  out_key3[3 * j + 1] = pk;
  out_key3[3 * j + 2] = zk;
  out_p[j] = D.hit_p[h];
}

}  // namespace mlm
