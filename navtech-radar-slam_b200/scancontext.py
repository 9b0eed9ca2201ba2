"""Host-side mirror of the reference's ScanContext operator API over the C ABI.

``SCManager`` keeps the method names, argument meaning and return conventions of
pgo/SC-A-LOAM/include/scancontext/Scancontext.h:57-122 (makeScancontext, make*key,
fastAlignUsingVkey, distDirectSC, distanceBtnScanContext, makeAndSaveScancontextAndKeys,
detectLoopClosureID, saveScancontextAndKeys, detectLoopClosureIDBetweenSession,
getConstRefRecentSCD, setSCdistThres, the public constants and the public data vectors) so
the parity tests read like calls into the reference.  All arithmetic happens in
libnrs_b200.so on the GPU; matrices are numpy float64 arrays of shape (20, 60) in Fortran
(column-major) order -- the memory layout of the reference's Eigen::MatrixXd.

``ShardedSCDatabase`` is the multi-GPU form north_star asks for: one process per GPU, each
rank holds a slice of the keyframes, every rank scores the same query against its slice
and the 16-byte per-shard minima are exchanged and merged with the exact lowest-index
tie-break the sequential scan would apply -- on GPUs inside the search kernel itself, over
peer memory (NVLink; nrs_sc_query_dev_sharded), otherwise with ONE collective
(torch.distributed all_gather: NCCL, or gloo in the CPU tests of the host logic).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import ScResult, check, f64, lib, ptr

REC_DTYPE = np.dtype([("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")])   # nrs_sc_result (16 bytes)

RINGS, SECTORS, DESC = 20, 60, 1200
SENTINEL = 10000000.0


def as_desc(m):
    """(20,60) column-major float64 view/copy of a descriptor given in any layout."""
    a = np.asarray(m, dtype=np.float64)
    if a.shape == (RINGS, SECTORS):
        return np.asfortranarray(a)
    if a.size == DESC:
        return np.asfortranarray(a.reshape(SECTORS, RINGS).T)
    raise ValueError(f"descriptor must be 20x60 or 1200 values, got shape {a.shape}")


def flat(m):
    """The 1200 doubles of a descriptor in reference (column-major) memory order."""
    return np.ascontiguousarray(as_desc(m).T).reshape(-1)


def best_of(results):
    """Exact merge of per-shard (dist, global_idx, shift): the strict-'<' sequential scan of
    SC/Scancontext.cpp:380-396 keeps the lowest index among equal distances; a shard with
    nothing found reports dist == 1e7 and never wins."""
    bd, bi, bs = SENTINEL, 0, 0
    found = False
    for d, i, s in results:
        if not (d < SENTINEL):
            continue
        if (not found) or d < bd or (d == bd and i < bi):
            bd, bi, bs, found = d, i, s, True
    return bd, bi, bs


class SCDatabase:
    """nrs_sc_db: device-resident mirror of SCManager::polarcontexts_ (Scancontext.h:111)."""

    def __init__(self, ctx, capacity_hint=1024):
        self.ctx = ctx
        self._h = C.c_void_p()
        check(lib().nrs_sc_db_create(ctx.handle, capacity_hint, C.byref(self._h)))
        ctx.adopt(self)

    def __len__(self):
        return check(lib().nrs_sc_db_size(self._h))

    def clear(self):
        check(lib().nrs_sc_db_clear(self._h))

    def append(self, descs):
        d = np.ascontiguousarray(descs, dtype=np.float64).reshape(-1, DESC)
        check(lib().nrs_sc_db_append(self._h, ptr(d), d.shape[0]))

    def append_cloud(self, xyz, lidar_height=2.0, max_radius=80.0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        out = np.empty(DESC, dtype=np.float64)
        check(lib().nrs_sc_db_append_cloud(self._h, ptr(xyz), xyz.shape[0], lidar_height,
                                           max_radius, ptr(out)))
        return out

    def read(self, first, count):
        out = np.empty((count, DESC), dtype=np.float64)
        check(lib().nrs_sc_db_read(self._h, first, count, ptr(out)))
        return out

    # -- peer-memory result exchange (keyframe-sharded, one process per GPU) -------------------------
    def xchg_create(self, world, rank):
        """-> this rank's 64-byte CUDA IPC handle (bytes)"""
        buf = (C.c_ubyte * 64)()
        check(lib().nrs_sc_xchg_create(self._h, world, rank, buf))
        return bytes(buf)

    def xchg_open(self, handles):
        """handles: the world's handles concatenated in rank order (world x 64 bytes)"""
        raw = bytes(handles)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        check(lib().nrs_sc_xchg_open(self._h, buf))

    def query_dev_sharded(self, query_dev_ptr, first, last, result_dev_ptr):
        """local slots [first,last); *result_dev = GLOBAL (dist, idx, shift) after the in-kernel exchange"""
        check(lib().nrs_sc_query_dev_sharded(self._h, C.c_void_p(query_dev_ptr), first, last, C.c_void_p(result_dev_ptr)))

    def tail_times(self):
        """%globaltimer ns stamps of the most recent search's last CTA (see nrs_sc_db_tail_times)"""
        t = (C.c_uint64 * 4)()
        check(lib().nrs_sc_db_tail_times(self._h, t))
        return [int(x) for x in t]

    def device_ptr(self):
        return lib().nrs_sc_db_device_ptr(self._h)

    def query(self, query, first=0, last=None, candidates=None):
        q = f64(query, DESC)
        last = len(self) if last is None else last
        res = ScResult()
        if candidates is None:
            check(lib().nrs_sc_query(self._h, ptr(q), first, last, None, 0, C.byref(res)))
        else:
            c = np.ascontiguousarray(candidates, dtype=np.int32)
            check(lib().nrs_sc_query(self._h, ptr(q), 0, 0, ptr(c), c.size, C.byref(res)))
        return res.dist, res.idx, res.shift

    def query_all(self, query, first=0, last=None):
        q = f64(query, DESC)
        last = len(self) if last is None else last
        n = max(0, last - first)
        dist = np.empty(n, dtype=np.float64)
        shift = np.empty(n, dtype=np.int32)
        res = ScResult()
        check(lib().nrs_sc_query_all(self._h, ptr(q), first, last, ptr(dist), ptr(shift),
                                     C.byref(res)))
        return dist, shift, (res.dist, res.idx, res.shift)

    def knn(self, ringkey, first=0, last=None, k=3):
        """nrs_sc_knn: the k nearest ring keys among keyframes [first,last) -> (ids int32, squared distances f32)"""
        key = np.ascontiguousarray(np.asarray(ringkey).reshape(-1), dtype=np.float32)
        assert key.size == RINGS
        last = len(self) if last is None else last
        idx = np.empty(k, dtype=np.int32)
        d2 = np.empty(k, dtype=np.float32)
        check(lib().nrs_sc_knn(self._h, ptr(key), first, last, k, ptr(idx), ptr(d2)))
        return idx, d2

    def detect_knn(self, query, first=0, last=None, k=3):
        """nrs_sc_detect_knn: ring key -> k nearest -> distanceBtnScanContext over them, all on the device.
        -> ((dist, idx, shift), candidate ids)"""
        q = f64(query, DESC)
        last = len(self) if last is None else last
        cand = np.empty(k, dtype=np.int32)
        res = ScResult()
        check(lib().nrs_sc_detect_knn(self._h, ptr(q), first, last, k, ptr(cand), C.byref(res)))
        return (res.dist, res.idx, res.shift), cand

    def detect_knn_batch(self, queries, first=0, last=None, k=3):
        """nrs_sc_detect_knn_batch: detect_knn for Q queries (Q,1200) in two launches
        -> ([(dist, idx, shift)] * Q, candidate ids (Q, k) int32)"""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DESC)
        last = len(self) if last is None else last
        cand = np.empty((q.shape[0], k), dtype=np.int32)
        res = (ScResult * q.shape[0])()
        check(lib().nrs_sc_detect_knn_batch(self._h, ptr(q), q.shape[0], first, last, k, ptr(cand), res))
        return [(r.dist, r.idx, r.shift) for r in res], cand

    def query_batch(self, queries, first=0, last=None, as_array=False):
        """nrs_sc_query_batch: Q queries (Q,1200) against keyframes [first,last) -> [(dist, idx, shift)] * Q,
        each exactly what query() returns; one kernel launch on a {0,c} database.  as_array: the Q records as
        one structured array (dist f8, idx i4, shift i4) instead of a list of tuples."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DESC)
        last = len(self) if last is None else last
        res = np.empty(q.shape[0], dtype=REC_DTYPE)
        check(lib().nrs_sc_query_batch(self._h, ptr(q), q.shape[0], first, last, ptr(res)))
        return res if as_array else [(float(r[0]), int(r[1]), int(r[2])) for r in res.tolist()]

    def query_batch_dev(self, queries_dev_ptr, Q, first, last, idx_offset, results_dev_ptr):
        check(lib().nrs_sc_query_batch_dev(self._h, C.c_void_p(queries_dev_ptr), Q, first, last, idx_offset,
                                           C.c_void_p(results_dev_ptr)))

    def query_batch_dev_sharded(self, queries_dev_ptr, Q, first, last, results_dev_ptr):
        """local slots [first,last); results_dev[q] = GLOBAL (dist, idx, shift) after ONE in-kernel exchange"""
        check(lib().nrs_sc_query_batch_dev_sharded(self._h, C.c_void_p(queries_dev_ptr), Q, first, last,
                                                   C.c_void_p(results_dev_ptr)))

    def is_binary(self):
        """True while every stored keyframe is {0,c}: searches run on the bit-mask kernel"""
        return bool(check(lib().nrs_sc_db_is_binary(self._h)))

    def set_search_mode(self, mode):
        """0 = auto (bit-mask kernel when applicable), 1 = always the general-valued f32-filter kernel"""
        check(lib().nrs_sc_db_set_search_mode(self._h, mode))

    def set_exchange_timeout_ms(self, ms):
        check(lib().nrs_sc_xchg_set_timeout_ms(self._h, ms))

    def query_dev(self, query_dev_ptr, first, last, idx_offset, result_dev_ptr):
        check(lib().nrs_sc_query_dev(self._h, C.c_void_p(query_dev_ptr), first, last, idx_offset,
                                     C.c_void_p(result_dev_ptr)))

    def close(self):
        if self._h:
            lib().nrs_sc_db_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SCManager:
    """Mirror of the reference's SCManager (Scancontext.h:57-122) on top of the GPU library."""

    # hyper parameters (Scancontext.h:83-104)
    LIDAR_HEIGHT = 2.0
    PC_NUM_RING = RINGS
    PC_NUM_SECTOR = SECTORS
    PC_MAX_RADIUS = 80.0
    PC_UNIT_SECTORANGLE = 360.0 / float(SECTORS)
    PC_UNIT_RINGGAP = 80.0 / float(RINGS)
    NUM_EXCLUDE_RECENT = 30
    NUM_CANDIDATES_FROM_TREE = 3
    SEARCH_RATIO = 0.1
    TREE_MAKING_PERIOD_ = 30

    def __init__(self, ctx=None, device=0, candidate_fn=None):
        """candidate_fn(keys_to_search (n,20) float32, query_key (20,) float32, k) -> ids.
        None (default) = exhaustive search over every searchable keyframe, which is what
        BASELINE.json's north_star specifies; "kdtree" = the reference's own flow on the device
        (ring-key k-NN with nanoflann's arithmetic, then the NUM_CANDIDATES_FROM_TREE candidates,
        nrs_sc_detect_knn); a callable supplies the candidate ids from the host (tests)."""
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.db = SCDatabase(self.ctx)
        self.SC_DIST_THRES = 0.2                       # Scancontext.h:99
        self.tree_making_period_conter = 0             # Scancontext.h:104
        # public data (Scancontext.h:110-116), kept populated on the host like the reference
        self.polarcontexts_timestamp_ = []
        self.polarcontexts_ = []
        self.polarcontext_invkeys_ = []
        self.polarcontext_vkeys_ = []
        self.polarcontext_invkeys_mat_ = []
        self.polarcontext_invkeys_to_search_ = []
        self.is_tree_batch_made = False
        self._batch_end = 0
        self._searchable_end = 0
        self._candidate_fn = candidate_fn

    # ---- descriptor construction -------------------------------------------------------------
    def makeScancontext(self, scan_down):
        xyz = np.ascontiguousarray(scan_down, dtype=np.float32).reshape(-1, 3)
        out = np.empty(DESC, dtype=np.float64)
        check(lib().nrs_sc_make(self.ctx.handle, ptr(xyz), xyz.shape[0], self.LIDAR_HEIGHT,
                                self.PC_MAX_RADIUS, ptr(out)))
        return as_desc(out)

    def _keys(self, desc):
        d = flat(desc)
        rk = np.empty(RINGS, dtype=np.float64)
        sk = np.empty(SECTORS, dtype=np.float64)
        check(lib().nrs_sc_keys(self.ctx.handle, ptr(d), ptr(rk), ptr(sk)))
        return rk.reshape(RINGS, 1), sk.reshape(1, SECTORS)

    def makeRingkeyFromScancontext(self, desc):
        return self._keys(desc)[0]

    def makeSectorkeyFromScancontext(self, desc):
        return self._keys(desc)[1]

    # ---- distances -----------------------------------------------------------------------------
    def fastAlignUsingVkey(self, vkey1, vkey2):
        a, b = f64(vkey1, SECTORS), f64(vkey2, SECTORS)
        s = C.c_int32()
        check(lib().nrs_sc_fast_align(self.ctx.handle, ptr(a), ptr(b), C.byref(s)))
        return s.value

    def distDirectSC(self, sc1, sc2):
        a, b = flat(sc1), flat(sc2)
        d = C.c_double()
        check(lib().nrs_sc_dist_direct(self.ctx.handle, ptr(a), ptr(b), C.byref(d)))
        return d.value

    def distanceBtnScanContext(self, sc1, sc2):
        a, b = flat(sc1), flat(sc2)
        d, s = C.c_double(), C.c_int32()
        check(lib().nrs_sc_pair(self.ctx.handle, ptr(a), ptr(b), C.byref(d), C.byref(s)))
        return d.value, s.value

    # ---- user-side API --------------------------------------------------------------------------
    def _save(self, sc):
        ringkey, sectorkey = self._keys(sc)
        self.polarcontexts_.append(sc)
        self.polarcontext_invkeys_.append(ringkey)
        self.polarcontext_vkeys_.append(sectorkey)
        self.polarcontext_invkeys_mat_.append(ringkey.reshape(-1).astype(np.float32))  # eig2stdvec

    def makeAndSaveScancontextAndKeys(self, scan_down):
        sc = as_desc(self.db.append_cloud(scan_down, self.LIDAR_HEIGHT, self.PC_MAX_RADIUS))
        self._save(sc)

    def saveScancontextAndKeys(self, scd):
        sc = as_desc(scd)
        self.db.append(flat(sc))
        self._save(sc)

    def getConstRefRecentSCD(self):
        return self.polarcontexts_[-1]

    def setSCdistThres(self, new_thres):
        self.SC_DIST_THRES = new_thres

    @staticmethod
    def _yaw(nn_align):
        degrees = np.float32(nn_align * (360.0 / float(SECTORS)))      # deg2rad(float), :17-20
        return float(np.float32(float(degrees) * math.pi / 180.0))

    def detectLoopClosureID(self):
        """Scancontext.cpp:331-422.  The searchable set follows the reference exactly: the
        keys present at the last tree rebuild minus the newest NUM_EXCLUDE_RECENT (rebuild on
        calls 0, 30, 60, ...).  Inside that set the default search is exhaustive."""
        loop_id = -1
        n = len(self.polarcontext_invkeys_mat_)
        if n < self.NUM_EXCLUDE_RECENT + 1:
            return loop_id, 0.0                                          # :341-345
        if self.tree_making_period_conter % self.TREE_MAKING_PERIOD_ == 0:   # :348-359
            self._searchable_end = n - self.NUM_EXCLUDE_RECENT
            self.polarcontext_invkeys_to_search_ = self.polarcontext_invkeys_mat_[: self._searchable_end]
        self.tree_making_period_conter += 1
        curr_desc = flat(self.polarcontexts_[-1])
        if self._candidate_fn is None:
            min_dist, nn_idx, nn_align = self.db.query(curr_desc, 0, self._searchable_end)
        elif self._candidate_fn == "kdtree":
            (min_dist, nn_idx, nn_align), self.last_candidates = self.db.detect_knn(
                curr_desc, 0, self._searchable_end, self.NUM_CANDIDATES_FROM_TREE)
        else:
            keys = np.stack(self.polarcontext_invkeys_to_search_).astype(np.float32)
            cand = self._candidate_fn(keys, self.polarcontext_invkeys_mat_[-1], self.NUM_CANDIDATES_FROM_TREE)
            min_dist, nn_idx, nn_align = self.db.query(curr_desc, candidates=cand)
        if min_dist < self.SC_DIST_THRES:                                  # :401
            loop_id = nn_idx
        self.last_min_dist, self.last_nn_idx = min_dist, nn_idx
        return loop_id, self._yaw(nn_align)                               # :417-420

    def detectLoopClosureIDBetweenSession(self, curr_key, curr_desc):
        """Scancontext.cpp:267-328.  The reference builds its batch tree ONCE, on the first call, over the
        keys present then (is_tree_batch_made): later calls search exactly those keyframes whatever has been
        saved since; no recent-exclusion.  Candidate ids the 3-NN search cannot fill stay 0 in the reference
        (zero-initialised vector, :292).  curr_key is only used in candidate mode."""
        if not self.is_tree_batch_made:
            self._batch_end = len(self.polarcontext_invkeys_mat_)
            self.polarcontext_invkeys_to_search_ = self.polarcontext_invkeys_mat_[: self._batch_end]
            self.is_tree_batch_made = True
        n = self._batch_end
        q = flat(curr_desc)
        k = self.NUM_CANDIDATES_FROM_TREE
        if n == 0:                                   # the reference would index an empty vector here (UB)
            return -1, self._yaw(0)
        if self._candidate_fn is None:
            min_dist, nn_idx, nn_align = self.db.query(q, 0, n)
        elif self._candidate_fn == "kdtree":
            # the reference searches the tree with the caller's curr_key (:297-300), whatever it is
            cand, _ = self.db.knn(curr_key, 0, n, k) if n > 0 else (np.full(k, -1, np.int32), None)
            cand = np.where(cand >= 0, cand, 0).astype(np.int32)
            self.last_candidates = cand
            min_dist, nn_idx, nn_align = self.db.query(q, candidates=cand) if n > 0 else (SENTINEL, 0, 0)
        else:
            keys = np.stack(self.polarcontext_invkeys_to_search_).astype(np.float32)
            cand = np.asarray(self._candidate_fn(keys, np.asarray(curr_key, dtype=np.float32), k), dtype=np.int32)
            cand = np.concatenate([np.where(cand >= 0, cand, 0), np.zeros(max(0, k - cand.size), np.int32)]).astype(np.int32)
            min_dist, nn_idx, nn_align = self.db.query(q, candidates=cand)
        loop_id = nn_idx if min_dist < self.SC_DIST_THRES else -1
        return loop_id, self._yaw(nn_align)


# -------------------------------------------------------------------------------------------------
# multi-GPU: keyframes sharded over ranks; per-shard records merged in-kernel over peer memory or by one collective
# -------------------------------------------------------------------------------------------------
def shard_owner(global_idx, world):
    """Round-robin ownership: keyframe g lives on rank g % world at local slot g // world, so a
    growing database stays balanced without ever moving a descriptor."""
    return global_idx % world, global_idx // world


def local_range(first, last, rank, world):
    """Local slots [lo, hi) of `rank` whose global index l*world+rank falls in [first,last)."""
    lo = max(0, -(-(first - rank) // world))
    hi = max(lo, -(-(last - rank) // world))
    return lo, hi


def enable_peer_exchange(dbs, rank, world, dev, group=None):
    """Set up the peer-memory result exchange (nrs_sc_xchg_*) on every database in `dbs` (same list
    length on every rank).  The IPC handles travel through one all_gather on the process group --
    plumbing, done once.  Returns True when every rank mapped every peer; False leaves the
    databases usable with the collective exchange."""
    import torch
    import torch.distributed as dist
    ok = True
    for db in dbs:
        mine = torch.frombuffer(bytearray(db.xchg_create(world, rank)), dtype=torch.uint8).to(dev)
        allh = torch.empty(64 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        try:
            db.xchg_open(allh.cpu().numpy().tobytes())
        except Exception:
            ok = False
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class ShardedSCDatabase:
    """Keyframe database sharded across the ranks of a torch.distributed group.

    Every rank calls append()/query() with identical arguments (SPMD).  Rank r stores
    keyframes g with g % world == r.  query() = local search on each rank's slice, then the
    exchange of the 16-byte (dist, idx, shift) records and their exact merge: on GPUs inside the
    search kernel over peer memory (exchange == "peer", nrs_sc_query_dev_sharded), else ONE
    all_gather + best_of() on the host (exchange == "nccl"; the CPU tests use gloo).
    ``local_search(query, lo, hi) -> (dist, local_idx, shift)`` defaults to the CUDA library;
    the CPU tests of this host logic inject a test double (gloo backend, no GPU).
    """

    RECORD = 16

    def __init__(self, rank, world, ctx=None, group=None, local_search=None, device=None, exchange="auto"):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.rank, self.world, self.group = rank, world, group
        self.n_total = 0
        self._local_n = 0
        self._ext_search = local_search
        if local_search is None:
            self.ctx = ctx if ctx is not None else _lib.Context(0 if device is None else device)
            self.db = SCDatabase(self.ctx)
            self.dev = torch.device("cuda", self.ctx.device)
            self._nrs_stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
            self._q_dev = torch.empty(DESC, dtype=torch.float64, device=self.dev)
            self._q_pin = torch.empty(DESC, dtype=torch.float64).pin_memory()
            self._rec_dev = torch.zeros(self.RECORD, dtype=torch.uint8, device=self.dev)
            self._all_dev = torch.zeros(self.RECORD * world, dtype=torch.uint8, device=self.dev)
            self._all_pin = torch.zeros(self.RECORD * world, dtype=torch.uint8).pin_memory()
            self._qb_dev = self._qb_pin = self._rb_dev = self._rb_pin = self._rb_all = None   # batch staging, lazily
            self.exchange = "none" if world == 1 else "nccl"
            if world > 1 and exchange in ("auto", "peer"):
                if enable_peer_exchange([self.db], rank, world, self.dev, group):
                    self.exchange = "peer"
                elif exchange == "peer":
                    raise RuntimeError("peer-memory exchange requested but CUDA IPC mapping failed on some rank")
        else:
            self.ctx = self.db = None
            self.dev = torch.device("cpu")
            self._store = []
            self.exchange = "collective"

    def __len__(self):
        return self.n_total

    def append(self, descs):
        d = np.ascontiguousarray(descs, dtype=np.float64).reshape(-1, DESC)
        mine = [i for i in range(d.shape[0]) if (self.n_total + i) % self.world == self.rank]
        if mine:
            if self.db is not None:
                self.db.append(d[mine])
            else:
                self._store.extend(d[mine])
            self._local_n += len(mine)
        self.n_total += d.shape[0]

    def local_descs(self):
        return np.stack(self._store) if self._store else np.zeros((0, DESC))

    # -- one search on this rank's slice -> 16-byte record tensor on self.dev ----------------------
    def _local_record(self, q, first, last):
        torch = self._torch
        lo, hi = local_range(first, last, self.rank, self.world)
        hi = min(hi, self._local_n)
        if self._ext_search is not None:
            dist_, lidx, shift = self._ext_search(q, lo, hi)
            rec = np.zeros(1, dtype=[("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")])
            rec["dist"], rec["idx"], rec["shift"] = dist_, lidx, shift
            return torch.from_numpy(rec.view(np.uint8).copy())
        self._q_pin.copy_(torch.from_numpy(q))
        with torch.cuda.stream(self._nrs_stream):
            self._q_dev.copy_(self._q_pin, non_blocking=True)
            self.db.query_dev(self._q_dev.data_ptr(), lo, hi, 0, self._rec_dev.data_ptr())
        return self._rec_dev

    def query(self, query, first=0, last=None):
        """Global (dist, idx, shift) over keyframes [first,last); identical on every rank."""
        torch, dist = self._torch, self._dist
        last = self.n_total if last is None else last
        q = f64(query, DESC)
        if self.world == 1 and self.db is not None:
            # one GPU: the plain C-ABI call (host query in, host result out), nothing to exchange
            return self.db.query(q, max(first, 0), max(first, min(last, self._local_n)))
        if self.exchange == "peer":
            # local search + in-kernel exchange over peer memory: the record that comes back is global
            lo, hi = local_range(first, last, self.rank, self.world)
            hi = min(hi, self._local_n)
            self._q_pin.copy_(torch.from_numpy(q))
            with torch.cuda.stream(self._nrs_stream):
                self._q_dev.copy_(self._q_pin, non_blocking=True)
                self.db.query_dev_sharded(self._q_dev.data_ptr(), lo, hi, self._rec_dev.data_ptr())
                pin = self._all_pin[: self.RECORD]
                pin.copy_(self._rec_dev, non_blocking=True)
            self._nrs_stream.synchronize()
            r = pin.numpy().view([("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")])[0]
            if int(r["idx"]) == -2:
                raise RuntimeError("sharded ScanContext query: a peer rank did not answer within the exchange timeout")
            return float(r["dist"]), int(r["idx"]), int(r["shift"])
        rec = self._local_record(q, first, last)
        if self.world == 1:
            gathered = rec
            if self.db is not None:
                torch.cuda.current_stream(self.dev).wait_stream(self._nrs_stream)
        elif self.db is not None:
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_stream(self._nrs_stream)
            dist.all_gather_into_tensor(self._all_dev, rec, group=self.group)
            gathered = self._all_dev
        else:
            outs = [torch.zeros(self.RECORD, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(outs, rec, group=self.group)
            gathered = torch.cat(outs)
        if self.db is not None:
            pin = self._all_pin[: gathered.numel()]
            pin.copy_(gathered, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            raw = pin.numpy()
        else:
            raw = gathered.numpy()
        recs = raw.view([("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")])
        n_r = recs.shape[0]
        base = self.rank if n_r == 1 else 0
        triples = [(float(recs["dist"][r]), int(recs["idx"][r]) * self.world + (r + base),
                    int(recs["shift"][r])) for r in range(n_r)]
        return best_of(triples)

    # -- batched search: Q queries, ONE exchange ------------------------------------------------------
    def _batch_buffers(self, Q):
        torch = self._torch
        if self._qb_dev is None or self._qb_dev.shape[0] < Q:
            cap = max(Q, 64)
            self._qb_dev = torch.empty((cap, DESC), dtype=torch.float64, device=self.dev)
            self._qb_pin = torch.empty((cap, DESC), dtype=torch.float64).pin_memory()
            self._rb_dev = torch.zeros(cap * self.RECORD, dtype=torch.uint8, device=self.dev)
            self._rb_pin = torch.zeros(cap * self.RECORD * max(1, self.world), dtype=torch.uint8).pin_memory()
            self._rb_all = torch.zeros(cap * self.RECORD * max(1, self.world), dtype=torch.uint8, device=self.dev)

    def query_batch(self, queries, first=0, last=None, as_array=False):
        """[(dist, idx, shift)] * Q over keyframes [first,last), each identical to query(); identical on every
        rank (as_array: one structured array (dist f8, idx i4, shift i4) instead of the list).  One GPU: nrs_sc_query_batch.  Sharded on GPUs: nrs_sc_query_batch_dev_sharded (one launch, the
        per-query records exchanged in-kernel over peer memory, merged once).  Otherwise one all_gather of
        the Q records per rank and best_of() per query."""
        torch, dist = self._torch, self._dist
        last = self.n_total if last is None else last
        qs = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, DESC)
        Q = qs.shape[0]
        rec_t = [("dist", "<f8"), ("idx", "<i4"), ("shift", "<i4")]
        if self.world == 1 and self.db is not None:
            return self.db.query_batch(qs, max(first, 0), max(first, min(last, self._local_n)), as_array)
        lo, hi = local_range(first, last, self.rank, self.world)
        hi = min(hi, self._local_n)
        if self.db is None:
            # CPU test double: per-query local search, one collective for the whole batch
            recs = np.zeros(Q, dtype=rec_t)
            for k in range(Q):
                recs["dist"][k], recs["idx"][k], recs["shift"][k] = self._ext_search(qs[k], lo, hi)
            mine = torch.from_numpy(recs.view(np.uint8).copy())
            outs = [torch.zeros(Q * self.RECORD, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(outs, mine, group=self.group)
            allr = [o.numpy().view(rec_t) for o in outs]
            return [best_of([(float(allr[r]["dist"][k]), int(allr[r]["idx"][k]) * self.world + r, int(allr[r]["shift"][k]))
                             for r in range(self.world)]) for k in range(Q)]
        if self.exchange == "peer":
            # nrs_sc_query_batch_sharded: the caller's buffer goes to the device as it is (DMA when it is pinned), one
            # launch with the in-kernel exchange, the Q records come back -- no staging copy on the host
            r = np.empty(Q, dtype=REC_DTYPE)
            check(lib().nrs_sc_query_batch_sharded(self.db._h, ptr(qs), Q, lo, hi, ptr(r)))
            if np.any(r["idx"] == -2):
                raise RuntimeError("sharded ScanContext batch: a peer rank did not answer within the exchange timeout")
            return r if as_array else [(float(x[0]), int(x[1]), int(x[2])) for x in r.tolist()]
        self._batch_buffers(Q)
        self._qb_pin[:Q].copy_(torch.from_numpy(qs))
        nb = Q * self.RECORD
        with torch.cuda.stream(self._nrs_stream):
            self._qb_dev[:Q].copy_(self._qb_pin[:Q], non_blocking=True)
            self.db.query_batch_dev(self._qb_dev.data_ptr(), Q, lo, hi, 0, self._rb_dev.data_ptr())
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self._nrs_stream)
        dist.all_gather_into_tensor(self._rb_all[: nb * self.world], self._rb_dev[:nb], group=self.group)
        self._rb_pin[: nb * self.world].copy_(self._rb_all[: nb * self.world], non_blocking=True)
        cur.synchronize()
        allr = self._rb_pin[: nb * self.world].numpy().view(rec_t).reshape(self.world, Q)
        return [best_of([(float(allr[r]["dist"][k]), int(allr[r]["idx"][k]) * self.world + r, int(allr[r]["shift"][k]))
                         for r in range(self.world)]) for k in range(Q)]
