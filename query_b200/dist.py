"""Multi-GPU plumbing (one process per GPU, torch.distributed): the Intermediate -> Final exchange of partial
group tables (execution/group_intermediate.go:56-104) and the global dictionary / statistics agreement that
makes string ranks and packed keys comparable across ranks.

Only the *movement* of records happens here (NCCL over NVLink via torch.distributed; gloo on CPU in tests);
producing the owner-bucketed records and merging them are CUDA kernels in libn1gpu.so
(k_count_owners / k_export_records / k_merge_records).

  small state (ungrouped, dense tables): all_gather of every rank's few records, merged on every rank
                                         (rank 0 reports);
  hash tables / DISTINCT sets:          all_to_all by owner = mix(group key) % world; each owner merges and
                                         finalises its share of the groups.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def row_range(nrows, r=None, w=None):
    """Contiguous range partition of the key-sorted document sequence (datastore/file/file.go:711-730 order)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    return nrows * r // w, nrows * (r + 1) // w


def exchange_by_owner(records: torch.Tensor, counts, words_per_record: int, group=None):
    """records: int64 tensor holding sum(counts) records of `words_per_record` words, bucketed by owner rank in rank
    order; counts[r] = records destined to rank r.  Returns (received records tensor, per-source counts)."""
    w = world()
    counts = [int(c) for c in counts]
    if w == 1:
        return records[: counts[0] * words_per_record], counts
    dev = records.device
    send_counts = torch.tensor(counts, dtype=torch.int64, device=dev)
    recv_counts = torch.empty(w, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    rc = [int(c) for c in recv_counts.tolist()]
    out = torch.empty(max(1, sum(rc)) * words_per_record, dtype=torch.int64, device=dev)
    dist.all_to_all_single(out[: sum(rc) * words_per_record], records[: sum(counts) * words_per_record].contiguous(),
                           output_split_sizes=[c * words_per_record for c in rc],
                           input_split_sizes=[c * words_per_record for c in counts], group=group)
    return out[: sum(rc) * words_per_record], rc


def gather_all(records: torch.Tensor, count: int, words_per_record: int, group=None):
    """Every rank contributes `count` records (counts may differ); every rank receives all of them."""
    w = world()
    if w == 1:
        return records[: count * words_per_record], [count]
    dev = records.device
    cnt = torch.tensor([count], dtype=torch.int64, device=dev)
    allc = torch.empty(w, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, cnt, group=group)
    cs = [int(c) for c in allc.tolist()]
    mx = max(1, max(cs))
    padded = torch.zeros(mx * words_per_record, dtype=torch.int64, device=dev)
    padded[: count * words_per_record] = records[: count * words_per_record]
    out = torch.empty(w * mx * words_per_record, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * mx * words_per_record: r * mx * words_per_record + cs[r] * words_per_record] for r in range(w)]
    return torch.cat(parts) if parts else out[:0], cs


def _merged_stats(allst, ndict):
    """the union of per-partition column statistics (n1gpu_table_stats_get layout)"""
    g = np.array(allst[0], dtype=np.int64)
    has = [s[1] != 0 for s in allst]
    g[0] = int(np.bitwise_or.reduce([s[0] for s in allst]))
    g[1] = int(any(has))
    mins = [s[2] for s, h in zip(allst, has) if h]
    maxs = [s[3] for s, h in zip(allst, has) if h]
    g[2] = min(mins) if mins else 0
    g[3] = max(maxs) if maxs else 0
    g[4] = int(any(s[4] != 0 for s in allst))
    g[5] = ndict
    g[7] = sum(s[7] for s in allst)  # MISSING / NULL rows of the whole keyspace
    return g


def agree_dictionaries_and_stats(table, group=None):
    """Before seal: merge every column's dictionary across ranks into the global sorted dictionary and take
    the union of the statistics, so that all ranks compile the same kernel and pack keys identically.
    One collective for everything (rows, raw dictionary blobs, statistics of all columns); the merge and the remapping of
    ranks are native (n1gpu_table_dict_merge: a kernel when the column already lives in HBM)."""
    w = world()
    if w == 1:
        table.set_global_rows(table.num_rows)
        return
    ncols = len(table.columns)
    mine = [int(table.num_rows)] + [table.dictionary_raw(c) + (table.stats(c).tolist(),) for c in range(ncols)]
    every = [None] * w
    dist.all_gather_object(every, mine, group=group)
    table.set_global_rows(sum(e[0] for e in every))  # exact bound for the overflow proofs (packed counters, one-word int sums)
    me = rank()
    for c in range(ncols):
        parts = [(e[1 + c][0], e[1 + c][1]) for e in every]
        same = all(len(o) == len(parts[0][1]) and b.size == parts[0][0].size and np.array_equal(o, parts[0][1]) and np.array_equal(b, parts[0][0])
                   for b, o in parts[1:])
        if not same:  # identical dictionaries (e.g. handed in with the columns): nothing to remap
            table.merge_dictionaries(c, [p for i, p in enumerate(parts) if i != me])
        ndict = len(parts[0][1]) - 1 if same else len(table.dictionary_raw(c)[1]) - 1
        table.set_stats(c, _merged_stats([e[1 + c][2] for e in every], ndict))


def agree_local(tables):
    """The same agreement between the partitions of one keyspace held by ONE process (several ranks driven in-process)."""
    total = sum(int(t.num_rows) for t in tables)
    for t in tables:
        t.set_global_rows(total)
    for c in range(len(tables[0].columns)):
        dicts = [t.dictionary(c) for t in tables]
        merged = sorted(set().union(*[set(d) for d in dicts]))
        allst = [t.stats(c).tolist() for t in tables]
        g = _merged_stats(allst, len(merged))
        for t, d in zip(tables, dicts):
            if merged != list(d):
                t.import_dictionary(c, merged)
            t.set_stats(c, g)


def make_mailbox(max_words=8192, group=None, arena_bytes=64 << 20):
    """Creates this rank's peer mailbox (+ arena for peer-readable direct-indexed group tables) and wires it to every
    other rank's (CUDA IPC handles exchanged through torch.distributed).  Returns None for a single rank."""
    from .api import Mailbox
    w = world()
    if w == 1:
        return None
    mb = Mailbox(w, rank(), max_words, arena_bytes)
    handles = [None] * w
    dist.all_gather_object(handles, mb.ipc_handle(), group=group)
    mb.open_peers(handles)
    dist.barrier(group=group)  # nobody pushes before every mailbox is mapped everywhere
    return mb


class _DevWords:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class DistributedQuery:
    """Runs a compiled query on this rank's row range and merges the partial groups across ranks.

    launch()/collect() pipeline small-state chains without host round trips: scan -> all_gather of the
    accumulator words (NCCL, stream-ordered) -> k_merge_words, all enqueued on `stream`; several steps can be in
    flight on different streams.  Hash-table / DISTINCT chains use the blocking owner exchange in collect()."""

    def __init__(self, query, group=None, stream=None, mailbox=None):
        self.q = query
        self.group = group
        self.stream = stream  # torch.cuda.Stream the query was bound to with set_stream (None: current stream)
        # slot == group key on every rank (one slot / dense table / direct-indexed HBM table): the merge is element-wise
        self.small = query.info["mode"] in ("ungrouped", "dense-shared-memory", "hbm-direct") and "distinct" not in " ".join(query.aggregates)
        # with a peer mailbox the merge of a small-state chain is fused into the scan: launch()/collect() only
        self.fused = bool(mailbox is not None and self.small and world() > 1 and query.info["mode"] != "hbm-direct")
        if world() > 1:
            # ranks whose kernels differ (a layout threshold crossed on one rank only) would enter different collectives and
            # merge slot-indexed words with hashed records: refuse that here, where it is an error message and not a hang
            # (one small all-reduce of a digest: min == max on every rank <=> all digests are equal)
            import hashlib
            text = "%s|%r|%s" % (query.info["mode"], tuple(query.word_ops()), query.kernel_source)
            h = int.from_bytes(hashlib.sha1(text.encode()).digest()[:7], "little")
            dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
            d = torch.tensor([h, -h], dtype=torch.int64, device=dev)
            dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
            hi, neg_lo = (int(x) for x in d.tolist())
            if hi != -neg_lo:
                raise RuntimeError("ranks compiled different kernels for one chain (declare the keyspace rows and agree the statistics "
                                   "before seal: agree_dictionaries_and_stats); this rank: mode %s, %d accumulator words"
                                   % (query.info["mode"], len(query.word_ops())))
        # A direct-indexed HBM table (megabytes, slot == key on every rank) lives in the mailbox arena: every rank folds and
        # finalises its slot range of all ranks' tables over NVLink.  A partitioned DISTINCT aggregation keeps its records
        # there: rank r aggregates and finalises partition range r.  Neither needs a collective or a replicated finalisation.
        self.peer = self.peer_part = False
        if mailbox is not None and world() > 1:
            query.set_mailbox(mailbox)
            self.peer_part = query.peer_mode == 2
            self.peer = query.peer_mode in (1, 2)
            if not (self.fused or self.peer):
                query.set_mailbox(None)
        if self.peer:
            self.small = False  # the result is owner-sharded, not replicated
        elif self.small and world() > 1 and stream is None and not self.fused:
            # the NCCL all_gather is ordered against torch's current stream: the scan must run on that stream too
            query.set_stream(torch.cuda.current_stream().cuda_stream)
        self._recs = None
        self._dents = None
        self._all = None
        self._view = None
        self._runs = None
        self._launched = False

    @property
    def replicated(self):
        """True when the merge leaves the complete result on EVERY rank (small state: element-wise merge), False when every
        group is finalised by exactly one owner rank (or there is one rank)."""
        return world() > 1 and self.small

    def describe(self):
        if world() == 1:
            return "none (one rank)"
        if self.peer_part:
            return ("owner-sharded and collective-free: (group, value) records partitioned by group range into a CUDA-IPC peer arena, flags per rank "
                    "and step, rank r aggregates partition range r from every rank's records over NVLink and finalises those groups")
        if self.peer:
            return ("owner-sharded and collective-free: direct-indexed tables in a CUDA-IPC peer arena, a release flag per rank and step, "
                    "rank r's k_finalize_groups folds slot range r of every rank's table over NVLink and finalises those groups")
        if self.fused:
            return "fused into nq_scan: peer stores over NVLink into every rank's mailbox + 1-block fold"
        if self.small and self.q.info["mode"] == "hbm-direct":
            return "in-place NCCL all-reduce of the direct-indexed table, one call per run of sum / min / max words, stream-ordered; replicated finalisation"
        if self.small:
            return "NCCL all_gather of the accumulator words + merge kernel, stream-ordered"
        return "owner-bucketed record / DISTINCT-entry export + NCCL all_to_all + merge kernel; every owner finalises its groups"

    def launch(self):
        q = self.q
        w = world()
        if w == 1 or self.peer:
            q.launch()   # peer: the flag that publishes this rank's table is enqueued behind the scan by the library
            self._launched = True
            return
        if not self.small:
            self._launched = True
            return
        q.launch()
        if self.fused:  # peer stores + mailbox merge are already enqueued by the library
            self._launched = True
            return
        if self._view is None:
            ptr, n = q.state_words()
            self._view = torch.as_tensor(_DevWords(ptr, n), device="cuda")
            self._runs = self._reduce_runs(n) if q.info["mode"] == "hbm-direct" else None
            if self._runs is None:
                self._all = torch.empty(w * n, dtype=torch.int64, device="cuda")
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _null()
        with ctx:
            if self._runs is not None:
                # megabytes of direct-indexed table: reduce in place, one NCCL all-reduce per run of words that combine
                # the same way (sum / min / max; NVLS reduces inside the switch) - every rank ends with the merged table
                for view, op in self._runs:
                    dist.all_reduce(view, op=op, group=self.group)
            else:
                dist.all_gather_into_tensor(self._all, self._view, group=self.group)  # stream-ordered after the scan
        if self._runs is None:
            q.merge_words(self._all.data_ptr(), w)
        self._launched = True

    def _reduce_runs(self, nwords):
        """[(tensor view over consecutive word planes, ReduceOp)] when every word's combine operation is one NCCL has
        (add u64 as int64 sum - two's complement; add f64; min / max i64), else None (all_gather + merge kernel)."""
        ops = self.q.word_ops()
        if not ops or nwords % len(ops):
            return None
        slots = nwords // len(ops)
        kind = {0: ("i", dist.ReduceOp.SUM), 1: ("f", dist.ReduceOp.SUM), 2: ("i", dist.ReduceOp.MIN), 3: ("i", dist.ReduceOp.MAX)}
        if any(o not in kind for o in ops):
            return None
        runs, w0 = [], 0
        for w in range(1, len(ops) + 1):
            if w == len(ops) or kind[ops[w]] != kind[ops[w0]]:
                view = self._view[w0 * slots: w * slots]
                dt, op = kind[ops[w0]]
                runs.append((view.view(torch.float64) if dt == "f" else view, op))
                w0 = w
        return runs

    def collect(self):
        self._launched = False
        if world() == 1 or self.small or self.peer:
            return self.q.collect()
        return self.execute()

    def _buffers(self, ng, nd, rw):
        need = max(1, ng) * rw
        if self._recs is None or self._recs.numel() < need:
            self._recs = torch.empty(need, dtype=torch.int64, device="cuda")  # no fill: nothing here may race the export kernels
        need = max(1, nd) * 2
        if self._dents is None or self._dents.numel() < need:
            self._dents = torch.empty(need, dtype=torch.int64, device="cuda")

    def execute(self):
        """Returns a Result holding this rank's share of the groups (small state: rank 0 holds all, others none)."""
        q = self.q
        w = world()
        if w == 1 or self.small or self.peer:
            self.launch()
            return self.q.collect()
        q.scan_partial()
        ng, nd, rw = q.partial_counts()
        self._buffers(ng, nd, rw)
        has_distinct = nd > 0 or "distinct" in " ".join(q.aggregates)
        if self.small and not has_distinct:
            counts, _ = q.partial_export(1, self._recs.data_ptr(), ng, self._dents.data_ptr(), nd)
            allrec, cs = gather_all(self._recs, int(counts[0]), rw, self.group)
            q.partial_reset()
            if rank() == 0:
                n = sum(cs)
                q.partial_import(allrec.data_ptr() if n else 0, n, 0, 0)
            return q.finalize()
        counts, dcounts = q.partial_export(w, self._recs.data_ptr(), ng, self._dents.data_ptr(), nd)
        got, rc = exchange_by_owner(self._recs, counts, rw, self.group)
        dgot, drc = exchange_by_owner(self._dents, dcounts, 2, self.group)
        torch.cuda.current_stream().synchronize()  # the import kernels run on the query's stream: records must have landed
        q.partial_reset()
        n, ndg = sum(rc), sum(drc)
        q.partial_import(got.data_ptr() if n else 0, n, dgot.data_ptr() if ndg else 0, ndg)
        return q.finalize()


def gather_results(operator, result, group=None):
    """Groups finalised by several ranks (every owner holds its share) -> ONE result on every rank, so that
    the operator's tail (HAVING / projection / ORDER BY / LIMIT need all groups) can run over it:
    Result.arrays() of every rank, string tables concatenated and the string payloads re-based, Operator.import_arrays."""
    kc, kv, ac, av, strings = result.arrays()
    w = world()
    if w == 1:
        return operator.import_arrays(kc, kv, ac, av, strings)
    parts = [None] * w
    dist.all_gather_object(parts, (kc, kv, ac, av, strings), group=group)
    all_strings, out = [], [[], [], [], []]
    for pkc, pkv, pac, pav, pstr in parts:
        base = len(all_strings)
        pkv, pav = pkv.copy(), pav.copy()
        pkv[pkc == 6] += base  # C_STRING payloads index this rank's string table
        pav[pac == 6] += base
        all_strings += list(pstr)
        for dst, a in zip(out, (pkc, pkv, pac, pav)):
            dst.append(a.reshape(-1))
    cat = [np.concatenate(a) if a else np.zeros(0) for a in out]
    return operator.import_arrays(cat[0], cat[1], cat[2], cat[3], all_strings)
