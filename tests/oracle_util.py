"""ctypes binding of the CPU oracle (oracle/_build/libgrb_oracle.so).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libgrb_oracle.so"))
        u64, u32, sz, vp, dbl = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p, C.c_double
        P = C.POINTER
        L.grbo_make_seed_pattern.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_uint, P(C.c_char_p)]
        L.grbo_calc_optimal_size.restype = u64
        L.grbo_calc_optimal_size.argtypes = [u64, C.c_uint, dbl]
        L.grbo_default_hash_universe.restype = u64
        L.grbo_default_hash_universe.argtypes = [u64, u64, u64]
        L.grbo_calc_phred_average.argtypes = [C.c_char_p, sz, P(u32), P(u32), P(dbl)]
        L.grbo_sum_phred.restype = dbl
        L.grbo_sum_phred.argtypes = [C.c_char_p, sz]
        L.grbo_hash_sequence.restype = sz
        L.grbo_hash_sequence.argtypes = [C.c_char_p, sz, P(C.c_char_p), C.c_uint, vp, sz]
        L.grbo_filter_new.restype = vp
        L.grbo_filter_new.argtypes = [u64, C.c_uint]
        L.grbo_filter_free.argtypes = [vp]
        L.grbo_filter_insert_bv.argtypes = [vp, vp, sz]
        L.grbo_filter_setup.restype = u64
        L.grbo_filter_setup.argtypes = [vp]
        L.grbo_filter_words.restype = P(u64)
        L.grbo_filter_words.argtypes = [vp, P(u64)]
        L.grbo_filter_rank.restype = u64
        L.grbo_filter_rank.argtypes = [vp, u64, P(C.c_int)]
        L.grbo_filter_get_id.restype = u32
        L.grbo_filter_get_id.argtypes = [vp, u64]
        L.grbo_filter_get_count.restype = u32
        L.grbo_filter_get_count.argtypes = [vp, u64]
        L.grbo_filter_set.argtypes = [vp, u64, u32, u32]
        L.grbo_filter_reset_ids.argtypes = [vp]
        L.grbo_query_tile.restype = u32
        L.grbo_query_tile.argtypes = [vp, vp, sz, P(u32), P(u32), vp, vp, u32, vp]
        L.grbo_insert_mibf.argtypes = [vp, vp, sz, u32]
        L.grbo_smooth_tiles.restype = sz
        L.grbo_smooth_tiles.argtypes = [sz, vp, vp, vp, vp, vp, u64]
        L.grbo_find_longest_stretch.argtypes = [vp, sz, P(C.c_int64), P(C.c_int64)]
        L.grbo_eval_flanks.argtypes = [C.c_int64, C.c_int64, vp, sz, P(u64), P(u64)]
        L.grbo_ntcard.restype = u64
        L.grbo_ntcard.argtypes = [C.c_char_p, P(C.c_char_p), C.c_uint, P(u64)]
        L.grbo_ntcard_sized.restype = u64
        L.grbo_ntcard_sized.argtypes = [C.c_char_p, P(C.c_char_p), C.c_uint, u64, P(u64)]
        _L = L
    return _L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _seed_arr(seeds):
    return (C.c_char_p * len(seeds))(*[s.encode() for s in seeds])


def make_seed_pattern(preset, k, w, h):
    bufs = [C.create_string_buffer(k + h + 2) for _ in range(h)]
    arr = (C.c_char_p * h)(*[C.cast(b, C.c_char_p) for b in bufs])
    lib().grbo_make_seed_pattern((preset or "").encode(), k, w, h, arr)
    return [b.value.decode() for b in bufs]


def calc_phred_average(qual: bytes):
    a, d = C.c_uint32(), C.c_uint32()
    s = (C.c_double * 2)()
    lib().grbo_calc_phred_average(qual, len(qual), C.byref(a), C.byref(d), s)
    return a.value, d.value, s[0], s[1]


def hash_sequence(seq: bytes, seeds):
    k, h = len(seeds[0]), len(seeds)
    frames = len(seq) - k + 1
    out = np.zeros(max(0, frames) * h, dtype=np.uint64)
    n = lib().grbo_hash_sequence(seq, len(seq), _seed_arr(seeds), h, _p(out), out.size)
    return out.reshape(-1, h)[:n]


class Filter:
    def __init__(self, bits, h):
        self.L = lib()
        self.h = h
        self.bits = bits
        self.f = self.L.grbo_filter_new(bits, h)

    def __del__(self):
        if getattr(self, "f", None):
            self.L.grbo_filter_free(self.f)
            self.f = None

    def insert_bv(self, hashes):
        a = np.ascontiguousarray(hashes, dtype=np.uint64).ravel()
        self.L.grbo_filter_insert_bv(self.f, _p(a), a.size)

    def setup(self):
        return self.L.grbo_filter_setup(self.f)

    def words(self):
        n = C.c_uint64()
        p = self.L.grbo_filter_words(self.f, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def rank(self, pos):
        b = C.c_int()
        r = self.L.grbo_filter_rank(self.f, int(pos), C.byref(b))
        return r, b.value

    def get(self, rank):
        return (self.L.grbo_filter_get_id(self.f, int(rank)),
                self.L.grbo_filter_get_count(self.f, int(rank)))

    def set(self, rank, id_, count):
        self.L.grbo_filter_set(self.f, int(rank), int(id_), int(count))

    def reset_ids(self):
        self.L.grbo_filter_reset_ids(self.f)

    def query_tile(self, hashes, cand_cap=64):
        a = np.ascontiguousarray(hashes, dtype=np.uint64)
        frames = a.shape[0]
        bi, bc = C.c_uint32(), C.c_uint32()
        ci = np.zeros(cand_cap, dtype=np.uint32)
        cc = np.zeros(cand_cap, dtype=np.uint32)
        cnt = np.zeros(3, dtype=np.uint64)
        n = self.L.grbo_query_tile(self.f, _p(a), frames, C.byref(bi), C.byref(bc), _p(ci), _p(cc),
                                   cand_cap, _p(cnt))
        return bi.value, bc.value, n, ci[:min(n, cand_cap)], cc[:min(n, cand_cap)], cnt

    def insert_mibf(self, hashes, id_):
        a = np.ascontiguousarray(hashes, dtype=np.uint64).ravel()
        self.L.grbo_insert_mibf(self.f, _p(a), a.size, int(id_))


def smooth_tiles(ids, assigned, cand_lists, threshold):
    n = len(ids)
    ids = np.array(ids, dtype=np.uint32)
    as_ = np.array(assigned, dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.uint32)
    ci, cc = [], []
    for i, cl in enumerate(cand_lists):
        for (a, b) in cl:
            ci.append(a)
            cc.append(b)
        off[i + 1] = len(ci)
    ci = np.array(ci + [0], dtype=np.uint32)
    cc = np.array(cc + [0], dtype=np.uint32)
    na = lib().grbo_smooth_tiles(n, _p(ids), _p(as_), _p(off), _p(ci), _p(cc), threshold)
    return ids, as_, na


def find_longest_stretch(assigned):
    a = np.array(assigned, dtype=np.uint8)
    s, e = C.c_int64(), C.c_int64()
    lib().grbo_find_longest_stretch(_p(a), len(a), C.byref(s), C.byref(e))
    return s.value, e.value


def eval_flanks(ls, le, ids):
    a = np.array(ids, dtype=np.uint32)
    ts, te = C.c_uint64(), C.c_uint64()
    g = lib().grbo_eval_flanks(ls, le, _p(a), len(a), C.byref(ts), C.byref(te))
    return bool(g), ts.value, te.value
