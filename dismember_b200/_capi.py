"""ctypes binding of libdismember_gpu.so -- the same symbols the JNI shim binds.

There is no fallback: if the shared library is missing or no CUDA device is
present, constructing an Engine raises.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMG_LIB") or os.path.join(_HERE, "libdismember_gpu.so")   # DMG_LIB: instrumented dev builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dismember_gpu.h")

DMG_OK, DMG_ERR_INVALID_ARG, DMG_ERR_CUDA, DMG_ERR_INDEX, DMG_ERR_STATE, DMG_ERR_UNSUPPORTED, DMG_ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6
DMG_F32, DMG_F64 = 0, 1


class DmgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[dmg {code}] {msg}")
        self.code = code


class DmgIndexError(DmgError, IndexError):
    """ArrayIndexOutOfBoundsException of LookupTable.embeddingLookup."""


class DmgArgumentError(DmgError, ValueError):
    """IllegalArgumentException / require(...)"""


def declared_symbols() -> list:
    """Every DMG_API function declared in include/dismember_gpu.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"DMG_API\s+[\w\s\*]+?\b(dmg_\w+)\s*\(", text)))


_lib = None


def load_library(build_if_missing: bool = True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise OSError(f"{LIB_PATH} not built (run python -m dismember_b200.build)")
        from . import build as _build
        _build.build()
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "dmg_create": [i32, C.POINTER(vp)],
        "dmg_destroy": [vp],
        "dmg_clone": [vp, C.POINTER(vp)],
        "dmg_set_stream": [vp, vp],
        "dmg_synchronize": [vp],
        "dmg_set_profiling": [vp, i32],
        "dmg_set_arithmetic": [vp, i32],
        "dmg_fast_stats": [vp, vp],
        "dmg_set_fast_tolerance": [vp, dbl],
        "dmg_wave_probe": [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp],
        "dmg_kernel_time": [vp, C.POINTER(dbl), C.POINTER(i64)],
        "dmg_load_tree_tdm": [vp, i32, i64, vp, vp, vp, i64, vp, vp, vp],
        "dmg_load_tree_complete": [vp, i32, i64, vp, vp],
        "dmg_load_din_weights": [vp, i32, i64, i32, i32, vp],
        "dmg_init_din_weights": [vp, i32, i64, i32, i32, u64],
        "dmg_download_din_weights": [vp, vp, i64],
        "dmg_otm_pseudo_targets": [vp, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp],
        "dmg_din_shape": [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        "dmg_tdm_retrieve": [vp, i32, vp, i32, i32, i32, vp, vp, i32, vp, vp, vp],
        "dmg_tdm_retrieve_dev": [vp, i32, vp, i32, i32, i32, vp, vp, vp],
        "dmg_tdm_retrieve_dev_sync": [vp, i32, vp, i32, i32, i32, vp, vp, vp],
        "dmg_otm_beam_search": [vp, i32, vp, i32, i32, vp, vp, vp],
        "dmg_otm_retrieve": [vp, i32, vp, i32, i32, i32, vp, vp, vp],
        "dmg_otm_beam_search_levels": [vp, i32, vp, i32, i32, vp, vp, vp],
        "dmg_score_pairs": [vp, i64, vp, vp, vp, i64, vp],
        "dmg_dr_load": [vp, i32, i32, i32, i32, i32, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp],
        "dmg_dr_load_paths": [vp, vp, vp],
        "dmg_dr_beam_search": [vp, i32, vp, i32, vp, vp, vp],
        "dmg_dr_retrieve": [vp, i32, vp, i32, i32, vp, vp, vp],
        "dmg_kmeans_tree": [vp, i32, i32, vp, i32, u64, vp],
        "dmg_set_sync_mode": [vp, i32],
        "dmg_dr_load_item_paths": [vp, i32, vp],
        "dmg_dr_init_synthetic": [vp, i32, i32, i32, i32, i32, i32, u64],
        "dmg_dr_train_step": [vp, i32, vp, vp, vp, i32, u64, dbl, i32, i32, i32, i32, vp, vp],
        "dmg_dr_download": [vp, i32, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp],
        "dmg_train_step": [vp, i64, vp, vp, vp, i64, vp, dbl, i32, vp],
        "dmg_train_step_dev": [vp, i64, vp, vp, vp, vp, dbl, i32, vp],
        "dmg_shard_train_step": [vp, i64, vp, vp, vp, i64, vp, dbl, i32, vp],
        "dmg_score_pairs_dev": [vp, i64, vp, vp, vp, vp],
        "dmg_din_gradients": [vp, i64, vp, vp, vp, i64, vp, vp, vp, i64],
        "dmg_tdm_sample_expand": [vp, i32, vp, vp, vp, i32, i32, i32, u64, vp, vp, vp, C.POINTER(i32)],
        "dmg_jtm_item_weights": [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dmg_jtm_assign_level": [vp, i32, vp, vp, i32, vp, i32, vp],
        "dmg_eval_metrics": [vp, i32, i32, vp, vp, vp, vp, vp],
        "dmg_load_deepfm_weights": [vp, i64, i32, i32, vp],
        "dmg_load_deepfm_weights_f64": [vp, i64, i32, i32, vp],
        "dmg_shard_unique_id": [vp, i32],
        "dmg_shard_init": [vp, i32, i32, vp],
        "dmg_shard_init_din_weights": [vp, i64, i32, i32, u64],
        "dmg_shard_load_din_weights": [vp, i64, i32, i32, vp],
        "dmg_shard_info": [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
        "dmg_shard_tdm_retrieve": [vp, i32, vp, i32, i32, i32, vp, vp, vp],
        "dmg_shard_jtm_item_weights": [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "dmg_shard_dr_load": [vp, i32, i32, i32, i32, i32, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp],
        "dmg_shard_dr_retrieve": [vp, i32, vp, i32, i32, vp, vp, vp],
        "dmg_dp_train_step": [vp, i64, vp, vp, vp, i64, vp, dbl, i32, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    L.dmg_last_error.argtypes = [vp]
    L.dmg_last_error.restype = C.c_char_p
    L.dmg_version.argtypes = []
    L.dmg_version.restype = C.c_char_p
    L.dmg_launch_count.argtypes = [vp]
    L.dmg_launch_count.restype = i64
    _lib = L
    return L


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def jtm_assign_level(parent_code, old_child, weights, max_assign) -> np.ndarray:
    """dmg_jtm_assign_level without a handle: TreeLearning.reBalance is host code inside the library, no device needed."""
    L = load_library()
    par, old = _i32(parent_code).ravel(), _i32(old_child).ravel()
    w = np.ascontiguousarray(weights, np.float32)
    out = np.empty(len(par), np.int32)
    rc = L.dmg_jtm_assign_level(None, len(par), _p(par), _p(old), w.shape[1], _p(w), int(max_assign), _p(out))
    if rc != DMG_OK:
        raise DmgArgumentError(rc, "dmg_jtm_assign_level: bad arguments")
    return out


class Engine:
    """One dmg handle = one CUDA device + stream (not thread-safe, like a model clone)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.dmg_create(device, C.byref(h))
        if rc != DMG_OK:
            raise DmgError(rc, (self.L.dmg_last_error(None) or b"").decode())
        self.h = h
        self.device = device
        self.din_dtype = None
        self.E = self.T = None
        self.rows = None

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc: int):
        if rc == DMG_OK:
            return
        msg = (self.L.dmg_last_error(self.h) or b"").decode()
        if rc == DMG_ERR_INDEX:
            raise DmgIndexError(rc, msg)
        if rc == DMG_ERR_INVALID_ARG:
            raise DmgArgumentError(rc, msg)
        raise DmgError(rc, msg)

    def clone(self) -> "Engine":
        """A handle that shares this engine's tree and weights read-only (own stream and scratch): one per host
        thread, like the reference's per-thread model clones (LocalOptimizer.scala:35-40).  Close clones first."""
        h = C.c_void_p()
        self._check(self.L.dmg_clone(self.h, C.byref(h)))
        e = Engine.__new__(Engine)
        e.L, e.h, e.device = self.L, h, self.device
        e.din_dtype, e.E, e.T, e.rows = self.din_dtype, self.E, self.T, self.rows
        e._deepfm = getattr(self, "_deepfm", False)
        if hasattr(self, "dr_shape"):
            e.dr_shape = self.dr_shape
        e._parent = self                                            # keeps the owner alive
        return e

    def close(self):
        if getattr(self, "h", None):
            self._check(self.L.dmg_destroy(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: Optional[int]):
        self._check(self.L.dmg_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        self._check(self.L.dmg_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.L.dmg_launch_count(self.h))

    def set_sync_mode(self, mode: str):
        """'spin' (default) | 'sleep': how the synchronous retrieval calls wait (dmg_set_sync_mode)."""
        self._check(self.L.dmg_set_sync_mode(self.h, {"spin": 0, "sleep": 1}[mode]))

    def set_arithmetic(self, mode: str):
        """'strict' | 'fast' (tensor-core scorer with certified cuts; same ids and logits)."""
        self._check(self.L.dmg_set_arithmetic(self.h, {"strict": 0, "fast": 1}[mode]))

    def set_fast_tolerance(self, tau: float):
        self._check(self.L.dmg_set_fast_tolerance(self.h, float(tau)))

    def wave_probe(self, item_seq, beam, level, use_mask=True):
        """Candidates of tree level `level` with their tensor-core scores and the bound eps (dmg_wave_probe)."""
        seq = _i32(item_seq).reshape(-1, self.T)
        B = len(seq)
        cap = max(((2 * beam + 7) // 8) * 8, 8)
        codes = np.empty((B, cap), np.int32)
        scores = np.empty((B, cap), np.float32)
        counts = np.empty(B, np.int32)
        eps = np.empty(B, np.float32)
        self._check(self.L.dmg_wave_probe(self.h, B, _p(seq), beam, int(use_mask), level, cap, _p(codes), _p(scores), _p(counts), _p(eps)))
        return codes, scores, counts, eps

    def din_shape(self):
        """-> (rows, embed_size, seq_len, dtype) of the loaded scorer"""
        r, e, t, d = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.L.dmg_din_shape(self.h, C.byref(r), C.byref(e), C.byref(t), C.byref(d)))
        return r.value, e.value, t.value, d.value

    def fast_stats(self):
        out = np.zeros(7, np.uint64)
        self._check(self.L.dmg_fast_stats(self.h, _p(out)))
        ratio = float(np.array([int(out[4]) & 0xFFFFFFFF], np.uint32).view(np.float32)[0])
        return {"cuts": int(out[0]), "cuts_rescored": int(out[1]), "rows_rescored": int(out[2]), "rows_fast": int(out[3]),
                "max_err_over_bound": ratio, "users_redone_strict": int(out[5]), "cuts_settled_in_place": int(out[6])}

    def set_profiling(self, on: bool):
        self._check(self.L.dmg_set_profiling(self.h, int(on)))

    def kernel_time(self):
        """-> (total ms of the beam-search kernel, launches) since the last call; synchronises."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self.L.dmg_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- index structures -------------------------------------------------------
    def load_tree_tdm(self, max_level, codes, node_ids, is_leaf, leaf_ids, leaf_codes, prob=None):
        codes, node_ids, leaf_ids, leaf_codes = _i32(codes), _i32(node_ids), _i32(leaf_ids), _i32(leaf_codes)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        pr = None if prob is None else np.ascontiguousarray(prob, np.float32)
        assert pr is None or len(pr) == len(codes)
        self._check(self.L.dmg_load_tree_tdm(self.h, int(max_level), len(codes), _p(codes), _p(node_ids), _p(is_leaf),
                                             len(leaf_ids), _p(leaf_ids), _p(leaf_codes), _p(pr)))

    def load_tree_complete(self, leaf_level, item_ids, leaf_ids):
        item_ids, leaf_ids = _i32(item_ids), _i32(leaf_ids)
        self._check(self.L.dmg_load_tree_complete(self.h, int(leaf_level), len(item_ids), _p(item_ids), _p(leaf_ids)))

    # -- weights ----------------------------------------------------------------
    def load_din_weights(self, params: np.ndarray, rows: int, E: int, T: int):
        if params.dtype == np.float32:
            dt = DMG_F32
        elif params.dtype == np.float64:
            dt = DMG_F64
        else:
            raise DmgArgumentError(DMG_ERR_INVALID_ARG, "DIN parameters must be float32 or float64")
        params = np.ascontiguousarray(params).ravel()
        n = rows * E + E * E + 2 * E * E + 2 * E + 1
        if params.size != n:
            raise DmgArgumentError(DMG_ERR_INVALID_ARG, f"compact DIN vector must hold {n} values, got {params.size}")
        self._check(self.L.dmg_load_din_weights(self.h, dt, rows, E, T, _p(params)))
        self.din_dtype, self.rows, self.E, self.T = params.dtype, rows, E, T
        self._deepfm = False

    def load_deepfm_weights(self, params: np.ndarray, rows: int, E: int, T: int):
        """DeepFM scorer: [emb | W1 (T+1)x(T+1)E | b1 | W2 | b2].  float32 = the TDM/JTM model (tdm/.../model/DeepFM.scala),
        float64 = OTM's DeepModel[Double] (otm/.../model/DeepFM.scala); the dtype of `params` decides."""
        params = np.asarray(params)
        dtype = np.dtype(np.float64 if params.dtype == np.float64 else np.float32)
        params = np.ascontiguousarray(params, dtype).ravel()
        n = rows * E + (T + 1) * (T + 1) * E + 2 * (T + 1) + 1
        if params.size != n:
            raise DmgArgumentError(DMG_ERR_INVALID_ARG, f"compact DeepFM vector must hold {n} values, got {params.size}")
        fn = self.L.dmg_load_deepfm_weights_f64 if dtype == np.float64 else self.L.dmg_load_deepfm_weights
        self._check(fn(self.h, rows, E, T, _p(params)))
        self.din_dtype, self.rows, self.E, self.T = dtype, rows, E, T
        self._deepfm = True

    def init_din_weights(self, dtype, rows: int, E: int, T: int, seed: int):
        dtype = np.dtype(dtype)
        dt = DMG_F32 if dtype == np.float32 else DMG_F64
        self._check(self.L.dmg_init_din_weights(self.h, dt, rows, E, T, seed))
        self.din_dtype, self.rows, self.E, self.T = dtype, rows, E, T
        self._deepfm = False

    def download_din_weights(self) -> np.ndarray:
        """Module.parameters() back from the device (DIN layout; a DeepFM model returns its own compact vector)."""
        n = self.rows * self.E + 3 * self.E * self.E + 2 * self.E + 1
        if getattr(self, "_deepfm", False):
            n = self.rows * self.E + (self.T + 1) * (self.T + 1) * self.E + 2 * (self.T + 1) + 1
        out = np.empty(n, self.din_dtype)
        self._check(self.L.dmg_download_din_weights(self.h, _p(out), n))
        return out

    def eval_metrics(self, rec_items, rec_counts, labels):
        """Metrics.computeMetrics per user -> [B, 3] (precision, recall, ndcg); labels: one sequence per user."""
        rec = _i32(rec_items)
        B, topk = rec.shape
        cnt = _i32(rec_counts).ravel()
        off = np.zeros(B + 1, np.int64)
        off[1:] = np.cumsum([len(l) for l in labels])
        flat = _i32(np.concatenate([np.asarray(l, np.int32) for l in labels])) if off[-1] else np.zeros(0, np.int32)
        out = np.empty((B, 3), np.float64)
        self._check(self.L.dmg_eval_metrics(self.h, B, topk, _p(rec), _p(cnt), _p(off), _p(flat), _p(out)))
        return out

    def jtm_assign_level(self, parent_code, old_child, weights, max_assign):
        """TreeLearning.reBalance for one level step (native host code in the library) -> new node per item."""
        par, old = _i32(parent_code).ravel(), _i32(old_child).ravel()
        w = np.ascontiguousarray(weights, np.float32)
        out = np.empty(len(par), np.int32)
        self._check(self.L.dmg_jtm_assign_level(self.h, len(par), _p(par), _p(old), w.shape[1], _p(w), int(max_assign), _p(out)))
        return out

    # -- node table sharded over the GPUs of one box (csrc/shard.cu) ---------------
    def shard_unique_id(self) -> bytes:
        buf = np.zeros(128, np.uint8)
        rc = self.L.dmg_shard_unique_id(_p(buf), 128)
        if rc != DMG_OK:
            raise DmgError(rc, "dmg_shard_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.tobytes()

    def shard_init(self, world: int, rank: int, unique_id: Optional[bytes] = None):
        uid = None if unique_id is None else np.frombuffer(unique_id, np.uint8).copy()
        self._check(self.L.dmg_shard_init(self.h, world, rank, _p(uid)))
        self.shard_world, self.shard_rank = world, rank

    def shard_init_din_weights(self, rows_global: int, E: int, T: int, seed: int):
        self._check(self.L.dmg_shard_init_din_weights(self.h, rows_global, E, T, seed))
        self.din_dtype, self.E, self.T = np.dtype(np.float32), E, T
        self.rows = self.shard_info()[0]

    def shard_load_din_weights(self, params: np.ndarray, rows_global: int, E: int, T: int):
        params = np.ascontiguousarray(params, np.float32).ravel()
        n = rows_global * E + 3 * E * E + 2 * E + 1
        if params.size != n:
            raise DmgArgumentError(DMG_ERR_INVALID_ARG, f"compact DIN vector must hold {n} values, got {params.size}")
        self._check(self.L.dmg_shard_load_din_weights(self.h, rows_global, E, T, _p(params)))
        self.din_dtype, self.E, self.T = np.dtype(np.float32), E, T
        self.rows = self.shard_info()[0]

    def shard_info(self):
        """-> (rows of the local table, rows of the whole table, candidates scored for other ranks so far)"""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.L.dmg_shard_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def shard_tdm_retrieve(self, item_seq, beam, topk, use_mask=True):
        """Collective: every rank calls it with its own B users (same B everywhere)."""
        seq = _i32(item_seq).reshape(-1, self.T)
        B = len(seq)
        items = np.empty((B, topk), np.int32)
        logits = np.empty((B, topk), np.float32)
        counts = np.empty(B, np.int32)
        self._check(self.L.dmg_shard_tdm_retrieve(self.h, B, _p(seq), beam, topk, int(use_mask), _p(items), _p(logits), _p(counts)))
        return items, logits, counts

    def shard_dr_load(self, num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b):
        """dmg_shard_dr_load: same (whole) tables as dr_load, only this rank's item range is uploaded."""
        c = lambda a: np.ascontiguousarray(a, np.float64)
        layer_w = [c(w) for w in layer_w]
        layer_b = [c(b) for b in layer_b]
        wp = (C.c_void_p * D)(*[w.ctypes.data for w in layer_w])
        bp = (C.c_void_p * D)(*[b.ctypes.data for b in layer_b])
        arrs = [c(layer_emb), c(rr_emb), c(rr_w), c(rr_b), c(sm_w), c(sm_b)]
        self._check(self.L.dmg_shard_dr_load(self.h, num_item, K, D, T, E, _p(arrs[0]), wp, bp, _p(arrs[1]), _p(arrs[2]),
                                             _p(arrs[3]), _p(arrs[4]), _p(arrs[5])))
        self.dr_shape = (num_item, K, D, T, E)

    def shard_dr_retrieve(self, seq, beam, topk):
        """Collective DeepRetrieval.recommend over the sharded item tables; every rank passes its own users."""
        _, K, D, T, E = self.dr_shape
        seq = _i32(seq).reshape(-1, T)
        B = len(seq)
        items = np.empty((B, topk), np.int32)
        sc = np.empty((B, topk), np.float64)
        counts = np.empty(B, np.int32)
        self._check(self.L.dmg_shard_dr_retrieve(self.h, B, _p(seq), beam, topk, _p(items), _p(sc), _p(counts)))
        return items, sc, counts

    def shard_jtm_item_weights(self, sample_off, sample_seq, parent_code, old_level, level, hierarchical=False, min_level=0,
                               use_mask=True):
        """Collective dmg_jtm_item_weights over the sharded table; every rank passes its own items."""
        off = np.ascontiguousarray(sample_off, np.int64)
        n_items = len(off) - 1
        seq = _i32(sample_seq).reshape(-1, self.T)
        par = _i32(parent_code).ravel()
        out = np.empty((n_items, 1 << (level - old_level)), np.float32)
        self._check(self.L.dmg_shard_jtm_item_weights(self.h, n_items, _p(off), _p(seq), _p(par), old_level, level,
                                                      int(hierarchical), int(min_level), int(use_mask), _p(out)))
        return out

    # -- retrieval --------------------------------------------------------------
    def tdm_retrieve(self, item_seq, beam, topk, use_mask=True, consumed_off=None, consumed=None, widen_beam=False):
        seq = _i32(item_seq).reshape(-1, self.T)
        B = len(seq)
        items = np.empty((B, topk), np.int32)
        logits = np.empty((B, topk), np.float32)
        counts = np.empty(B, np.int32)
        co = None if consumed_off is None else np.ascontiguousarray(consumed_off, np.int64)
        cc = None if consumed is None else _i32(consumed)
        self._check(self.L.dmg_tdm_retrieve(self.h, B, _p(seq), beam, topk, int(use_mask), _p(co), _p(cc),
                                            int(widen_beam), _p(items), _p(logits), _p(counts)))
        return items, logits, counts

    def tdm_retrieve_dev(self, B, d_seq_ptr, beam, topk, use_mask, d_items_ptr, d_logits_ptr, d_counts_ptr):
        """Device-pointer form (ints from tensor.data_ptr()); asynchronous on the handle's stream."""
        vp = C.c_void_p
        self._check(self.L.dmg_tdm_retrieve_dev(self.h, B, vp(d_seq_ptr), beam, topk, int(use_mask), vp(d_items_ptr),
                                                vp(d_logits_ptr), vp(d_counts_ptr)))

    def tdm_retrieve_dev_sync(self, B, d_seq_ptr, beam, topk, use_mask, d_items_ptr, d_logits_ptr, d_counts_ptr):
        """Device-pointer form that returns with the results complete (strict redo launched only when a batch needs it)."""
        vp = C.c_void_p
        self._check(self.L.dmg_tdm_retrieve_dev_sync(self.h, B, vp(d_seq_ptr), beam, topk, int(use_mask), vp(d_items_ptr),
                                                     vp(d_logits_ptr), vp(d_counts_ptr)))

    def otm_beam_search(self, leaf_seq, beam, use_mask=True):
        seq = _i32(leaf_seq).reshape(-1, self.T)
        B = len(seq)
        s = int(beam).bit_length() - 1
        width = 2 * max(beam, 1 << s)
        ids = np.empty((B, width), np.int32)
        sc = np.empty((B, width), np.float64)
        counts = np.empty(B, np.int32)
        self._check(self.L.dmg_otm_beam_search(self.h, B, _p(seq), beam, int(use_mask), _p(ids), _p(sc), _p(counts)))
        return ids, sc, counts

    def otm_beam_search_levels(self, leaf_seq, beam, leaf_level, use_mask=True):
        seq = _i32(leaf_seq).reshape(-1, self.T)
        B = len(seq)
        s = int(beam).bit_length() - 1
        width = 2 * max(beam, 1 << s)
        n_lvl = max(leaf_level - s, 0)
        ids = np.empty((B, n_lvl, width), np.int32)
        sc = np.empty((B, n_lvl, width), np.float64)
        counts = np.empty((B, n_lvl), np.int32)
        self._check(self.L.dmg_otm_beam_search_levels(self.h, B, _p(seq), beam, int(use_mask), _p(ids), _p(sc), _p(counts)))
        return ids, sc, counts

    def otm_retrieve(self, leaf_seq, beam, topk, use_mask=True):
        seq = _i32(leaf_seq).reshape(-1, self.T)
        B = len(seq)
        items = np.empty((B, topk), np.int32)
        sc = np.empty((B, topk), np.float64)
        counts = np.empty(B, np.int32)
        self._check(self.L.dmg_otm_retrieve(self.h, B, _p(seq), beam, topk, int(use_mask), _p(items), _p(sc), _p(counts)))
        return items, sc, counts

    def score_pairs(self, node, seq, mask_flat=None):
        node = _i32(node).ravel()
        seq = _i32(seq).reshape(len(node), self.T)
        out = np.empty(len(node), self.din_dtype)
        m = None if mask_flat is None else _i32(mask_flat).ravel()
        self._check(self.L.dmg_score_pairs(self.h, len(node), _p(node), _p(seq), _p(m), 0 if m is None else len(m), _p(out)))
        return out

    def score_pairs_dev(self, n, d_node_ptr, d_seq_ptr, d_mask_ptr, d_out_ptr):
        """Device-pointer form of score_pairs (ints from tensor.data_ptr(); d_mask_ptr = n x T mask bytes or 0); asynchronous."""
        vp = C.c_void_p
        self._check(self.L.dmg_score_pairs_dev(self.h, int(n), vp(d_node_ptr), vp(d_seq_ptr), vp(d_mask_ptr or None), vp(d_out_ptr)))

    # -- Deep Retrieval ------------------------------------------------------------
    def dr_load(self, num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b):
        c = lambda a: np.ascontiguousarray(a, np.float64)
        layer_w = [c(w) for w in layer_w]
        layer_b = [c(b) for b in layer_b]
        wp = (C.c_void_p * D)(*[w.ctypes.data for w in layer_w])
        bp = (C.c_void_p * D)(*[b.ctypes.data for b in layer_b])
        arrs = [c(layer_emb), c(rr_emb), c(rr_w), c(rr_b), c(sm_w), c(sm_b)]
        self._check(self.L.dmg_dr_load(self.h, num_item, K, D, T, E, _p(arrs[0]), wp, bp, _p(arrs[1]), _p(arrs[2]),
                                       _p(arrs[3]), _p(arrs[4]), _p(arrs[5])))
        self.dr_shape = (num_item, K, D, T, E)

    def dr_init_synthetic(self, num_item, K, D, T, E, J=2, seed=0):
        """synthetic model + path CSR generated on the device (whole tables, or this rank's item range after shard_init)"""
        self._check(self.L.dmg_dr_init_synthetic(self.h, num_item, K, D, T, E, J, int(seed)))
        self.dr_shape = (num_item, K, D, T, E)

    def dr_load_paths(self, path_off, path_items):
        po = np.ascontiguousarray(path_off, np.int64)
        pi = _i32(path_items)
        self._check(self.L.dmg_dr_load_paths(self.h, _p(po), _p(pi)))

    def dr_beam_search(self, seq, beam):
        _, K, D, T, E = self.dr_shape
        seq = _i32(seq).reshape(-1, T)
        B = len(seq)
        paths = np.empty((B, beam, D), np.int32)
        probs = np.empty((B, beam), np.float64)
        counts = np.empty(B, np.int32)
        self._check(self.L.dmg_dr_beam_search(self.h, B, _p(seq), beam, _p(paths), _p(probs), _p(counts)))
        return paths, probs, counts

    def dr_retrieve(self, seq, beam, topk):
        _, K, D, T, E = self.dr_shape
        seq = _i32(seq).reshape(-1, T)
        B = len(seq)
        items = np.empty((B, topk), np.int32)
        sc = np.empty((B, topk), np.float64)
        counts = np.empty(B, np.int32)
        self._check(self.L.dmg_dr_retrieve(self.h, B, _p(seq), beam, topk, _p(items), _p(sc), _p(counts)))
        return items, sc, counts

    def kmeans_tree(self, embeddings, iters, seed=0):
        """RecursiveCluster.run (kmeans) -> node code per point"""
        emb = np.ascontiguousarray(embeddings, np.float64)
        codes = np.full(len(emb), -1, np.int32)
        self._check(self.L.dmg_kmeans_tree(self.h, emb.shape[0], emb.shape[1], _p(emb), int(iters), int(seed), _p(codes)))
        return codes

    def dr_load_item_paths(self, item_paths):
        """itemPathMapping as [num_item, P, D] node indices"""
        ip = _i32(item_paths)
        num_item, K, D, T, E = self.dr_shape
        ip = ip.reshape(num_item, -1, D)
        self._check(self.L.dmg_dr_load_item_paths(self.h, ip.shape[1], _p(ip)))

    def dr_train_step(self, seq, target, lr, step_t, rerank_step_t=None, sampled=None, num_sampled=0, seed=0, parallelism=1, apply=True):
        """one mini-batch iteration of the Deep Retrieval LocalOptimizer -> (layer losses [D], rerank loss)"""
        num_item, K, D, T, E = self.dr_shape
        seq = _i32(seq).reshape(-1, T)
        tg = _i32(target).ravel()
        rerank_step_t = step_t if rerank_step_t is None else int(rerank_step_t)
        sp = None
        if sampled is not None:
            sp = _i32(sampled).reshape(len(seq), -1)
            num_sampled = sp.shape[1] - 1
        loss = np.zeros(D, np.float64)
        rloss = np.zeros(1, np.float64)
        self._check(self.L.dmg_dr_train_step(self.h, len(seq), _p(seq), _p(tg), _p(sp), int(num_sampled), int(seed), float(lr), int(step_t),
                                             rerank_step_t, int(parallelism), int(bool(apply)), _p(loss), _p(rloss)))
        return loss, float(rloss[0])

    def dr_download(self, gradients=False):
        """-> dict(layer_emb, layer_w[D], layer_b[D], rr_emb, rr_w, rr_b, sm_w, sm_b): parameters, or the gradients of the last apply=False step"""
        num_item, K, D, T, E = self.dr_shape
        out = {"layer_emb": np.empty((num_item + K * (D - 1), E)), "layer_w": [np.empty((K, (T + d) * E)) for d in range(D)],
               "layer_b": [np.empty(K) for _ in range(D)], "rr_emb": np.empty((num_item, E)), "rr_w": np.empty((E, T * E)),
               "rr_b": np.empty(E), "sm_w": np.empty((num_item, E)), "sm_b": np.empty(num_item)}
        wp = (C.c_void_p * D)(*[w.ctypes.data for w in out["layer_w"]])
        bp = (C.c_void_p * D)(*[b.ctypes.data for b in out["layer_b"]])
        self._check(self.L.dmg_dr_download(self.h, int(bool(gradients)), _p(out["layer_emb"]), wp, bp, _p(out["rr_emb"]), _p(out["rr_w"]),
                                           _p(out["rr_b"]), _p(out["sm_w"]), _p(out["sm_b"])))
        return out

    # -- training / JTM ---------------------------------------------------------------
    def din_gradients(self, node, seq, mask_flat, labels):
        node = _i32(node).ravel()
        seq = _i32(seq).reshape(len(node), self.T)
        labels = np.ascontiguousarray(labels, self.din_dtype).ravel()
        m = None if mask_flat is None else _i32(mask_flat).ravel()
        n = self.rows * self.E + 3 * self.E * self.E + 2 * self.E + 1
        if getattr(self, "_deepfm", False):
            n = self.rows * self.E + (self.T + 1) * (self.T + 1) * self.E + 2 * (self.T + 1) + 1
        grad = np.empty(n, self.din_dtype)
        loss = np.zeros(1, self.din_dtype)
        self._check(self.L.dmg_din_gradients(self.h, len(node), _p(node), _p(seq), _p(m), 0 if m is None else len(m),
                                             _p(labels), _p(loss), _p(grad), n))
        return grad, loss[0]

    def train_step(self, node, seq, mask_flat, labels, lr, step_t):
        node = _i32(node).ravel()
        seq = _i32(seq).reshape(len(node), self.T)
        labels = np.ascontiguousarray(labels, self.din_dtype).ravel()
        m = None if mask_flat is None else _i32(mask_flat).ravel()
        loss = np.zeros(1, self.din_dtype)
        self._check(self.L.dmg_train_step(self.h, len(node), _p(node), _p(seq), _p(m), 0 if m is None else len(m),
                                          _p(labels), float(lr), int(step_t), _p(loss)))
        return loss[0]

    def train_step_dev(self, rows, d_node_ptr, d_seq_ptr, d_mask_ptr, d_labels_ptr, lr, step_t, d_loss_ptr):
        """Device-pointer form of train_step; asynchronous on the handle's stream (index errors surface at synchronize())."""
        vp = C.c_void_p
        self._check(self.L.dmg_train_step_dev(self.h, int(rows), vp(d_node_ptr), vp(d_seq_ptr), vp(d_mask_ptr or None), vp(d_labels_ptr),
                                              float(lr), int(step_t), vp(d_loss_ptr)))

    def dp_train_step(self, node, seq, mask_flat, labels, lr, step_t):
        """Collective data-parallel step (dmg_dp_train_step): this rank's rows, gradients averaged over the ranks."""
        node = _i32(node).ravel()
        seq = _i32(seq).reshape(len(node), self.T)
        labels = np.ascontiguousarray(labels, self.din_dtype).ravel()
        m = None if mask_flat is None else _i32(mask_flat).ravel()
        loss = np.zeros(1, self.din_dtype)
        self._check(self.L.dmg_dp_train_step(self.h, len(node), _p(node), _p(seq), _p(m), 0 if m is None else len(m),
                                             _p(labels), float(lr), int(step_t), _p(loss)))
        return loss[0]

    def shard_train_step(self, node, seq, mask_flat, labels, lr, step_t):
        """Collective training step on the sharded table (dmg_shard_train_step): this rank's rows with GLOBAL node codes."""
        node = _i32(node).ravel()
        seq = _i32(seq).reshape(len(node), self.T)
        labels = np.ascontiguousarray(labels, np.float32).ravel()
        m = None if mask_flat is None else _i32(mask_flat).ravel()
        loss = np.zeros(1, np.float32)
        self._check(self.L.dmg_shard_train_step(self.h, len(node), _p(node), _p(seq), _p(m), 0 if m is None else len(m),
                                                _p(labels), float(lr), int(step_t), _p(loss)))
        return loss[0]

    def otm_pseudo_targets(self, leaf_seq, target_off, targets, leaf_level, start_level, use_mask=True, M=None):
        """OTMTree.optimalPseudoTargets on the device -> (ids [n_lvl, B, M], vals, counts [n_lvl, B]), levels start_level + 1 .. leaf_level"""
        seq = _i32(leaf_seq).reshape(-1, self.T)
        B = len(seq)
        off = np.ascontiguousarray(target_off, np.int64)
        tg = _i32(targets).ravel()
        M = int(M or max(1, int(np.diff(off).max())))
        n_lvl = leaf_level - start_level
        ids = np.empty((n_lvl, B, M), np.int32)
        vals = np.empty((n_lvl, B, M), np.float64)
        cnt = np.empty((n_lvl, B), np.int32)
        self._check(self.L.dmg_otm_pseudo_targets(self.h, B, _p(seq), _p(off), _p(tg), start_level, int(use_mask), M, _p(ids), _p(vals), _p(cnt)))
        return ids, vals, cnt

    def tdm_sample_expand(self, target_items, item_seq, layer_neg, start_level, seed, with_prob=False, tolerance=20):
        tg = _i32(target_items).ravel()
        seq = _i32(item_seq).reshape(len(tg), self.T)
        neg = _i32(layer_neg).ravel()
        layer_sum = int(sum(1 + int(x) for x in neg[start_level:]))
        rows = len(tg) * layer_sum
        node = np.empty(rows, np.int32)
        oseq = np.empty((rows, self.T), np.int32)
        lab = np.empty(rows, np.float32)
        n = C.c_int32()
        self._check(self.L.dmg_tdm_sample_expand(self.h, len(tg), _p(tg), _p(seq), _p(neg), start_level, int(with_prob), int(tolerance), seed, _p(node),
                                                 _p(oseq), _p(lab), C.byref(n)))
        assert n.value == rows
        return node, oseq, lab

    def jtm_item_weights(self, sample_off, sample_seq, parent_code, old_level, level, hierarchical=False, min_level=0,
                         use_mask=True):
        off = np.ascontiguousarray(sample_off, np.int64)
        n_items = len(off) - 1
        seq = _i32(sample_seq).reshape(-1, self.T)
        par = _i32(parent_code).ravel()
        out = np.empty((n_items, 1 << (level - old_level)), np.float32)
        self._check(self.L.dmg_jtm_item_weights(self.h, n_items, _p(off), _p(seq), _p(par), old_level, level,
                                                int(hierarchical), int(min_level), int(use_mask), _p(out)))
        return out
