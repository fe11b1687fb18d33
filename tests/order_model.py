"""Pure-numpy model of libstdc++'s unordered_map iteration order used by the
CPU tests to validate the data-parallel scheme the CUDA path implements (order_kernels.cuh):
given the distinct keys in first-insert order and the bucket count at frame start, return the
iteration order and the final bucket count."""
import numpy as np

BUCKET_CHAIN = [1, 13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933, 351061,
                712697, 1447153, 2938679, 5967347, 12117689, 24607243, 49969847, 101473717]


def vector_hash(keys):
    """VectorHasher, reference include/map_awareness.h:31-41 (int arithmetic, wraps)"""
    keys = np.asarray(keys, dtype=np.int64)
    h = np.full(keys.shape[0], 3, dtype=np.int64)
    for c in range(3):
        hu = h & 0xFFFFFFFF
        hs = np.where(hu >= 2 ** 31, hu - 2 ** 32, hu)  # as signed int
        term = ((keys[:, c] & 0xFFFFFFFF) + 0x9E3779B9 + ((hu << 6) & 0xFFFFFFFF) + ((hs >> 2) & 0xFFFFFFFF)) & 0xFFFFFFFF
        h = (hu ^ term) & 0xFFFFFFFF
    return np.where(h >= 2 ** 31, h - 2 ** 32, h).astype(np.int64)


def buckets(keys, B):
    h = vector_hash(keys)
    hu = h.astype(np.uint64)  # sign-extended two's complement of the int hash as size_t
    return (hu % np.uint64(B)).astype(np.int64)


def order(seq_keys, B):
    """iteration order of keys inserted in sequence seq_keys into a table with B buckets (no rehash)"""
    n = seq_keys.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    b = buckets(seq_keys, B)
    act = np.full(B, n, dtype=np.int64)
    np.minimum.at(act, b, np.arange(n))
    a = act[b]
    return np.lexsort((-np.arange(n), -a))  # descending activation, then descending position


def chain_next(B):
    return BUCKET_CHAIN[BUCKET_CHAIN.index(B) + 1]


def iteration_order(insert_seq, B0):
    """returns (keys in iteration order, final bucket count)"""
    seq = np.asarray(insert_seq).copy()
    n = seq.shape[0]
    B = B0
    if n > 0 and B == 1:
        B = 13
    while n > B:
        perm = order(seq[:B], B)
        seq[:B] = seq[:B][perm]
        B = chain_next(B)
    return seq[order(seq, B)], B
