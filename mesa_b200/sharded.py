"""Row-sharded Schelling on several GPUs: BASELINE.json config 5 (65536^2 on 8 x B200), SURVEY section 8e.

One process per GPU.  Rank k owns the contiguous row block `sharding.block_bounds(rows, world, k)` of
every per-cell array (type int8, happy uint8, agent_id int32, 16-byte relocation slots).  All of them
live in torch symmetric memory, i.e. they are peer-mapped over NVLink / NVSwitch, so

* `assign_state` is the single-GPU stencil kernel whose halo pointers are the neighbours' own edge
  rows (read in place, no exchange buffers), followed by a 1-element all-reduce of `model.happy`;
* `shuffle_do("step")` runs the single-GPU relocation kernels unchanged: a probe is a uniform draw over
  the WHOLE grid (grid.py:303-306), the kernels load the owner's slot, claim it with an NVLink
  atomicMin and finally write the arrival into the owner's layers.  The host only all-reduces the
  per-pass "changed" flag and runs the small collectives of the argwhere fallback
  (gather the few fallback movers, all-gather per-rank empty counts).

Agent ids are global (row-major over the whole grid, schelling/model.py:72-81), priorities and draws
are keyed on them, so the trajectory is identical for every GPU count -- and equal to the reference's.
The reference has no counterpart for any of this (single process).
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _native as N
from . import ops
from .sharding import block_bounds

_FB_BATCH = 2048     # kFbBatch of csrc/relocate.cu
_FB_BLOCKS = 65536 + 1024   # kFbBlocksMax + kFbSupersMax rows of kFbBatch words (blocks, then super-blocks)
_OV_CAP = 4096       # kOvCap
_NONE64 = -1  # 0xffff_ffff_ffff_ffff as int64


class _Phase:
    def __init__(self, owner, name):
        self.o, self.name = owner, name
        self.on = os.environ.get("MB_SHARD_PROFILE") is not None

    def __enter__(self):
        if self.on:
            import time

            torch.cuda.synchronize(self.o.device)
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if self.on:
            import time

            torch.cuda.synchronize(self.o.device)
            prof = self.o.__dict__.setdefault("profile", {})
            rec = prof.setdefault(self.name, [0, 0.0])
            rec[0] += 1
            rec[1] += (time.perf_counter() - self.t0) * 1e3
        return False


def _u64_sort_key(t: torch.Tensor) -> torch.Tensor:
    """int64 bit patterns of uint64 values -> keys whose signed order is the unsigned order."""
    return t ^ torch.iinfo(torch.int64).min


class ShardedSchelling:
    """Schelling model state + step on a row-sharded grid (mirror of schelling/model.py:31-93)."""

    def __init__(self, rows_total: int, cols: int, local_types, homophily: float = 0.4, radius: int = 1,
                 seed: int = 0, device=None, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        if not dist.is_initialized():
            raise RuntimeError("ShardedSchelling needs an initialised torch.distributed process group (nccl)")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.rows_total, self.cols = int(rows_total), int(cols)
        self.homophily, self.radius, self.seed = float(homophily), int(radius), int(seed)
        self.lo, self.hi = block_bounds(self.rows_total, self.world, self.rank)
        self.rows = self.hi - self.lo
        if self.rows < self.radius:
            raise ValueError("row block thinner than the neighbourhood radius")
        local_types = torch.as_tensor(local_types).to(torch.int8)
        if tuple(local_types.shape) != (self.rows, self.cols):
            raise ValueError(f"rank {self.rank} expects a [{self.rows}, {self.cols}] block of the type layer")
        max_rows = -(-self.rows_total // self.world)

        def symmetric(shape, dtype):
            t = symm_mem.empty(shape, dtype=dtype, device=self.device)
            return t, symm_mem.rendezvous(t, self.group)

        self._type_buf, self._type_h = symmetric((max_rows, self.cols), torch.int8)
        self._happy_buf, self._happy_h = symmetric((max_rows, self.cols), torch.uint8)
        self._id_buf, self._id_h = symmetric((max_rows, self.cols), torch.int32)
        self._slot_buf, self._slot_h = symmetric((max_rows * self.cols, 2), torch.int64)
        self.type = self._type_buf[: self.rows]
        self.happy = self._happy_buf[: self.rows]
        self.agent_id = self._id_buf[: self.rows]
        self.type.copy_(local_types.to(self.device))
        self.happy.zero_()

        # global agent ids: row-major over the whole grid
        occupied = self.type.reshape(-1) >= 0
        self.n_local = int(occupied.sum().item())
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        counts[self.rank] = self.n_local
        dist.all_reduce(counts, group=self.group)
        self.n_agents = int(counts.sum().item())
        if self.n_agents >= 2**32 - 1:
            raise ValueError("agent ids must fit 32 bits")
        first = int(counts[: self.rank].sum().item())
        ids = torch.zeros(self.rows * self.cols, dtype=torch.int64, device=self.device)
        ids[occupied] = torch.arange(first, first + self.n_local, dtype=torch.int64, device=self.device)
        self.agent_id.copy_(ids.reshape(self.rows, self.cols).to(torch.int32))  # uint32 bit pattern

        # bitmap of the agents that stay put in a step, one section per rank: mb_reloc_collect fills this rank's section,
        # an all-gather replicates them, and the probe kernels test a LOCAL, L2-resident bit before touching a (mostly
        # remote) 16-byte slot -- more than half of all probes on a 0.8-dense grid stop there
        self._bits_stride = -(-(max_rows * self.cols) // 32)
        self._bits = torch.zeros((self.world, self._bits_stride), dtype=torch.int32, device=self.device)
        self.shard = N.Shard()
        self.shard.world, self.shard.rank = self.world, self.rank
        self.shard.rows_total, self.shard.cols = self.rows_total, self.cols
        if os.environ.get("MB_RELOC_NOBITS") is None:
            self.shard.static_bits, self.shard.bits_stride = self._bits.data_ptr(), self._bits_stride
        for r in range(self.world):
            self.shard.slots[r] = self._slot_h.buffer_ptrs[r]
            self.shard.type[r] = self._type_h.buffer_ptrs[r]
            self.shard.happy[r] = self._happy_h.buffer_ptrs[r]
            self.shard.agent_id[r] = self._id_h.buffer_ptrs[r]
        self._shard_ref = ctypes.byref(self.shard)

        # neighbours' edge rows, read in place by the stencil kernel (clamped grid: none outside)
        self._halo_top = self._halo_bot = None
        if self.rank > 0:
            ulo, uhi = block_bounds(self.rows_total, self.world, self.rank - 1)
            up = self._type_h.get_buffer(self.rank - 1, (max_rows, self.cols), torch.int8)
            self._halo_top = up[uhi - ulo - self.radius: uhi - ulo]
        if self.rank < self.world - 1:
            down = self._type_h.get_buffer(self.rank + 1, (max_rows, self.cols), torch.int8)
            self._halo_bot = down[: self.radius]

        # agents migrate between ranks: a block can hold up to one agent per cell
        self.max_movers = max(self.rows * self.cols, 1)
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_reloc_workspace_bytes(self.max_movers, ctypes.byref(nbytes)))
        self._ws_bytes = int(nbytes.value)
        self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
        # event loop of the fallback cascade (csrc/relocate.cu): every rank's block-count table is peer-mapped, the loop
        # itself runs on rank 0 and reaches all slot / count tables over NVLink
        self._cnt_buf, self._cnt_h = symmetric((_FB_BATCH * _FB_BLOCKS,), torch.int32)
        self._cnt_ptrs = (ctypes.c_void_p * self.world)(*[self._cnt_h.buffer_ptrs[r] for r in range(self.world)])
        self._ov = torch.zeros(2 * _OV_CAP + 4, dtype=torch.int64, device=self.device)   # times | cells | out4
        self.last_events = [0, 0, 0, 0]      # overrides, bail reason, fallback events, displaced events (summed over buckets)
        self._table = ops.happy_threshold_table(self.homophily, self.radius)
        self._count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._host7 = (ctypes.c_uint64 * 7)()
        self.epoch = 0
        self.happy_count = 0
        self.last_movers = self.last_passes = 0
        self._stream = lambda: N.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().mb_reloc_slots_init(self._shard_ref, self._stream()))
        self._barrier()
        self.assign_state()

    # ------------------------------------------------------------------ plumbing
    def _phase(self, name: str):
        """MB_SHARD_PROFILE=1: wall time per phase of move_unhappy (device-synchronised on both sides) in `self.profile`."""
        return _Phase(self, name)

    def _barrier(self):
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)

    def _counters(self):
        N.check(N.lib().mb_reloc_counters(N.ptr(self._ws), self._ws_bytes, self.max_movers, self._host7, self._stream()))
        return [int(x) for x in self._host7]

    _REL_CAP = 4096   # kRelCap of csrc/relocate.cu

    def _pass_flags(self):
        """(fallback movers per rank, released cells per rank, table changed anywhere, full pass required) -- the counters
        of every rank, all-reduced; the collective also orders every rank's kernels before the next phase."""
        c = self._counters()
        w = self.world
        flags = torch.zeros(2 * w + 2, dtype=torch.int64, device=self.device)
        flags[self.rank], flags[w + self.rank], flags[2 * w], flags[2 * w + 1] = c[2], c[5], c[3], c[6]
        dist.all_reduce(flags, group=self.group)
        flags = flags.tolist()
        return flags[:w], flags[w:2 * w], flags[2 * w] != 0, flags[2 * w + 1] != 0

    def _take_releases(self, nrel_all):
        """Start of an incremental iteration: the union of every rank's release write set (cells whose registered arrival
        left, time of the leaver); clears the local set and the changed flag."""
        width = max(max(nrel_all), 1)
        mine = torch.zeros((2, width), dtype=torch.int64, device=self.device)
        N.check(N.lib().mb_reloc_release_take(N.ptr(self._ws), self._ws_bytes, self.max_movers, N.ptr(mine[0]),
                                              N.ptr(mine[1]), nrel_all[self.rank], self._stream()))
        if sum(nrel_all) == 0:
            return mine[0], mine[1], 0
        everyone = torch.empty((self.world, 2, width), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        cells = torch.cat([everyone[r, 0, :n] for r, n in enumerate(nrel_all)]).contiguous()
        times = torch.cat([everyone[r, 1, :n] for r, n in enumerate(nrel_all)]).contiguous()
        return cells, times, int(cells.numel())

    def _allreduce(self, values, op=dist.ReduceOp.SUM):
        t = torch.tensor(values, dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=op, group=self.group)
        return t.tolist()

    # ------------------------------------------------------------------ do("assign_state")
    def assign_state(self) -> int:
        """agents.py:24-42 for every agent of the block; returns the global model.happy."""
        ops.schelling_happy(self.type, self.homophily, self.radius, False, out=self.happy, count=self._count,
                            halo_top=self._halo_top, halo_bot=self._halo_bot, table=self._table)
        dist.all_reduce(self._count, group=self.group)
        self.happy_count = int(self._count.item())
        return self.happy_count

    # ------------------------------------------------------------------ shuffle_do("step")
    def move_unhappy(self) -> int:
        lib, ws, wb, mm = N.lib(), N.ptr(self._ws), self._ws_bytes, self.max_movers
        seed, epoch = self.seed & 0xFFFFFFFFFFFFFFFF, self.epoch & 0xFFFFFFFF
        self.epoch += 1
        with torch.cuda.device(self.device):
            with self._phase("collect"):
                N.check(lib.mb_reloc_collect(self._shard_ref, seed, epoch, ws, wb, mm, self._stream()))
                movers, occupied = self._counters()[:2]
                tot_movers, tot_occupied = self._allreduce([movers, occupied])   # also orders collect before any pass
            with self._phase("bitmap all_gather"):
                if self.world > 1 and self.shard.static_bits:
                    dist.all_gather_into_tensor(self._bits, self._bits[self.rank].clone(), group=self.group)
            self.last_movers, self.last_passes = tot_movers, 0
            self.last_events = [0, 0, 0, 0]
            if tot_movers == 0:
                return 0
            nempty = self.rows_total * self.cols - tot_occupied
            if nempty <= 0:
                raise ValueError("Grid is completely full. No empty cells available. Cannot select a random empty cell.")
            # rank-ordered buckets (earliest priorities first, each to convergence), ~16 M movers per GPU each
            buckets = 1
            while buckets < 256 and tot_movers // (buckets * self.world) > (16 << 20):
                buckets <<= 1
            forced = int(os.environ.get("MB_RELOC_BUCKETS", "0") or 0)      # the parity suite forces the bucketed path
            if 1 <= forced <= 256 and forced & (forced - 1) == 0:
                buckets = forced
            width = (1 << 64) // buckets
            # several buckets: the local mover indices are grouped by bucket once, a pass then walks only its bucket's list
            # instead of scanning every mover for the 1 / buckets that take part (csrc/relocate.cu, mb_reloc_bucket_lists)
            offsets = None
            if buckets > 1:
                counts = (ctypes.c_uint64 * buckets)()
                with self._phase("bucket lists"):
                    N.check(lib.mb_reloc_bucket_lists(movers, buckets, counts, ws, wb, mm, self._stream()))
                offsets = [0]
                for c in counts:
                    offsets.append(offsets[-1] + int(c))

            def run_pass(b, mode):
                if offsets is None:
                    N.check(lib.mb_reloc_pass(self._shard_ref, movers, seed, epoch, t_lo, t_hi, mode, ws, wb, mm, self._stream()))
                else:
                    N.check(lib.mb_reloc_pass_list(self._shard_ref, offsets[b], offsets[b + 1] - offsets[b], seed, epoch, t_lo, t_hi,
                                                   mode, ws, wb, mm, self._stream()))

            for b in range(buckets):
                t_lo = b * width
                t_hi = (t_lo + width - 1) if b + 1 < buckets else (1 << 64) - 1
                # 1. ordinary movers: FULL pass + RESUME passes until nothing changes anywhere (fallback movers only listed:
                #    no release can occur); 2. the fallback movers and what they displace, in time order, by the event loop on
                #    rank 0; 3. the generic iteration below confirms with a FULL iteration (csrc/relocate.cu).
                prefix_ok = os.environ.get("MB_RELOC_GENERIC") is None
                settled = False      # the bucket reached its fixed point by construction (csrc/relocate.cu, same rule)
                if prefix_ok:
                    with self._phase("FULL pass"):
                        run_pass(b, 0)
                    self.last_passes += 1
                    guard = 0
                    while True:
                        with self._phase("flags (D2H + all_reduce)"):
                            nfb_all, nrel_all, changed, force = self._pass_flags()
                        if force or sum(nrel_all) > 0:
                            prefix_ok = False
                            break
                        if guard > 0 and not changed:
                            break
                        guard += 1
                        with self._phase("RESUME pass"):
                            N.check(lib.mb_reloc_release_take(ws, wb, mm, None, None, 0, self._stream()))
                            run_pass(b, 1)
                        self.last_passes += 1
                    settled = prefix_ok and sum(nfb_all) == 0
                    if prefix_ok and 0 < sum(nfb_all) <= _FB_BATCH:
                        with self._phase("event phase"):
                            settled = self._event_phase(movers, nfb_all, nempty, seed, epoch, t_lo, t_hi)
                if settled and os.environ.get("MB_RELOC_CONFIRM", "0") in ("", "0"):
                    continue             # no confirming FULL iteration: claims only lowered `arr`, nothing was released
                full, nrel_all = True, [0] * self.world
                while True:
                    self.last_passes += 1
                    if full:
                        run_pass(b, 0)
                    else:
                        cells, times, total = self._take_releases(nrel_all)
                        if total > 0:
                            N.check(lib.mb_reloc_release_scan(self._shard_ref, movers, seed, epoch, t_lo, t_hi, N.ptr(cells),
                                                              N.ptr(times), total, ws, wb, mm, self._stream()))
                        run_pass(b, 1)
                    nfb_all, _, _, _ = self._pass_flags()       # also: every rank's pass has finished past this point
                    if sum(nfb_all) > 0:
                        self._serve_fallback(nfb_all[self.rank], nfb_all, nempty, seed, epoch)
                    _, nrel_all, changed, force = self._pass_flags()
                    was_full = full
                    if not (changed or force or sum(nrel_all) > 0):
                        if was_full:
                            break                                   # a full iteration without any change: the bucket is final
                        full = True                                 # confirm the incremental result
                        continue
                    full = force or sum(nrel_all) > self._REL_CAP
            with self._phase("vacate + arrive"):
                N.check(lib.mb_reloc_apply(self._shard_ref, 0, movers, None, ws, wb, mm, self._stream()))
                self._barrier()      # every origin is empty before any arrival is written
                N.check(lib.mb_reloc_apply(self._shard_ref, 1, movers, None, ws, wb, mm, self._stream()))
                self._barrier()
        return tot_movers

    def _gather_fallback(self, nfb_all: list[int], nempty: int, seed: int, epoch: int):
        """(time, k, mover) of every rank's fallback movers sorted by time + the rank each one lives on."""
        lib, ws, wb, mm = N.lib(), N.ptr(self._ws), self._ws_bytes, self.max_movers
        dev, world, nfb = self.device, self.world, nfb_all[self.rank]
        width = max(nfb_all)
        mine = torch.zeros((width, 3), dtype=torch.int64, device=dev)     # (time, k, mover) uint64 bit patterns
        if nfb > 0:
            t = torch.empty(nfb, dtype=torch.int64, device=dev)
            k = torch.empty(nfb, dtype=torch.int64, device=dev)
            m = torch.empty(nfb, dtype=torch.int32, device=dev)
            N.check(lib.mb_reloc_fb_export(self._shard_ref, 0, nfb, nempty, seed, epoch, N.ptr(t), N.ptr(k), N.ptr(m),
                                           ws, wb, mm, self._stream()))
            mine[:nfb, 0], mine[:nfb, 1], mine[:nfb, 2] = t, k, m.to(torch.int64) & 0xFFFFFFFF
        everyone = torch.empty((world, width, 3), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        rows, homes = [], []
        for r, n in enumerate(nfb_all):
            rows.append(everyone[r, :n])
            homes.append(torch.full((n,), r, dtype=torch.int64, device=dev))
        entries, home = torch.cat(rows), torch.cat(homes)
        order = ops.argsort_stable(_u64_sort_key(entries[:, 0]))
        return entries[order], home[order]

    def _event_phase(self, movers: int, nfb_all: list[int], nempty: int, seed: int, epoch: int, t_lo: int, t_hi: int) -> bool:
        """The fallback cascade of a converged bucket: block counts on every rank, the time-ordered event loop on rank 0
        (all slot / count tables peer-mapped), the changed picks broadcast back as overrides."""
        lib, ws, wb, mm = N.lib(), N.ptr(self._ws), self._ws_bytes, self.max_movers
        dev, world = self.device, self.world
        with self._phase("  ev: gather fallback movers"):
            entries, _ = self._gather_fallback(nfb_all, nempty, seed, epoch)
        count = int(entries.shape[0])
        times, ks = entries[:, 0].contiguous(), entries[:, 1].contiguous()
        totals = torch.zeros(_FB_BATCH, dtype=torch.int32, device=dev)
        with self._phase("  ev: block counts + all_gather"):
            N.check(lib.mb_reloc_fb_count_into(self._shard_ref, N.ptr(times), count, N.ptr(totals), N.ptr(self._cnt_buf), self._stream()))
            all_totals = torch.empty((world, _FB_BATCH), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(all_totals, totals, group=self.group)      # also: every rank's table is counted
        ov = self._ov
        with self._phase("  ev: event loop (rank 0) + broadcast"):
            if self.rank == 0:
                N.check(lib.mb_reloc_event_loop(self._shard_ref, self._cnt_ptrs, N.ptr(all_totals), N.ptr(times), N.ptr(ks), count,
                                                N.ptr(ov[:_OV_CAP]), N.ptr(ov[_OV_CAP:2 * _OV_CAP]), N.ptr(ov[2 * _OV_CAP:]),
                                                seed, epoch, self._stream()))
            dist.broadcast(ov, dist.get_global_rank(self.group, 0) if self.group is not dist.group.WORLD else 0, group=self.group)
        out4 = ov[2 * _OV_CAP:].tolist()
        e = self.last_events
        self.last_events = [e[0] + int(out4[0]), max(e[1], int(out4[1])), e[2] + int(out4[2]), e[3] + int(out4[3])]
        N.check(lib.mb_reloc_apply_overrides(movers, t_lo, t_hi, N.ptr(ov[:_OV_CAP]), N.ptr(ov[_OV_CAP:2 * _OV_CAP]), int(out4[0]),
                                             ws, wb, mm, self._stream()))
        return int(out4[1]) == 0     # bail reason 0: every event was processed in time order

    def _serve_fallback(self, nfb: int, nfb_all: list[int], nempty: int, seed: int, epoch: int) -> None:
        """grid.py:308-316 for the movers that missed all 50 probes, across ranks."""
        lib, ws, wb, mm = N.lib(), N.ptr(self._ws), self._ws_bytes, self.max_movers
        dev, world = self.device, self.world
        entries, home = self._gather_fallback(nfb_all, nempty, seed, epoch)
        for first in range(0, entries.shape[0], _FB_BATCH):
            batch = entries[first:first + _FB_BATCH]
            count = batch.shape[0]
            times = batch[:, 0].contiguous()
            totals = torch.zeros(count, dtype=torch.int32, device=dev)
            N.check(lib.mb_reloc_fb_count(self._shard_ref, N.ptr(times), count, N.ptr(totals), ws, wb, mm, self._stream()))
            all_totals = torch.empty((world, count), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(all_totals, totals, group=self.group)
            cum = torch.cumsum(all_totals.to(torch.int64), dim=0)            # empties in ranks 0..r at each time
            kk = batch[:, 1]
            owner = (cum <= kk.unsqueeze(0)).sum(dim=0)                        # first rank whose cumulative count exceeds k
            before = torch.where(owner > 0, cum[(owner - 1).clamp(min=0), torch.arange(count, device=dev)], 0)
            k_local = torch.where(owner == self.rank, kk - before, torch.full_like(kk, _NONE64)).contiguous()
            result = torch.zeros(count, dtype=torch.int64, device=dev)
            N.check(lib.mb_reloc_fb_pick(self._shard_ref, N.ptr(times), N.ptr(k_local), count, N.ptr(result),
                                         ws, wb, mm, self._stream()))
            dist.all_reduce(result, group=self.group)                          # exactly one rank wrote each entry
            movers32 = torch.where(home[first:first + count] == self.rank, batch[:, 2],
                                   torch.full_like(kk, -1)).to(torch.int32).contiguous()   # -1 = 0xffffffff: not mine
            N.check(lib.mb_reloc_fb_commit(self._shard_ref, N.ptr(times), N.ptr(movers32), N.ptr(result),
                                           count, ws, wb, mm, self._stream()))

    # ------------------------------------------------------------------ model step
    def step(self) -> int:
        """schelling/model.py:87-93: shuffle_do("step") then do("assign_state"); returns model.happy."""
        self.move_unhappy()
        return self.assign_state()

    def gather_types(self) -> torch.Tensor | None:
        """Whole type layer on rank 0 (tests / readback)."""
        max_rows = self._type_buf.shape[0]
        out = torch.empty((self.world, max_rows, self.cols), dtype=torch.int8, device=self.device)
        dist.all_gather_into_tensor(out, self._type_buf.contiguous(), group=self.group)
        blocks = []
        for r in range(self.world):
            lo, hi = block_bounds(self.rows_total, self.world, r)
            blocks.append(out[r, : hi - lo])
        return torch.cat(blocks)

    def gather_layer(self, buf: torch.Tensor) -> torch.Tensor:
        max_rows = buf.shape[0]
        out = torch.empty((self.world,) + tuple(buf.shape), dtype=buf.dtype, device=self.device)
        dist.all_gather_into_tensor(out, buf.contiguous(), group=self.group)
        return torch.cat([out[r, : block_bounds(self.rows_total, self.world, r)[1] - block_bounds(self.rows_total, self.world, r)[0]]
                          for r in range(self.world)])


class ShardedBoidsNCCL:
    """First version of the domain decomposition: torch-level packing + NCCL point-to-point (kept as the reference
    implementation of the protocol and as a fallback; ShardedBoids below is the peer-memory version).

    BoidFlockers on a continuous torus sharded by spatial domain (BASELINE.json config 4).

    Rank k owns the slab  y in [k, k+1) * height / world  and the boids inside it.  Per step:
      1. halo: boids within `vision` of a slab edge are copied to the neighbouring rank (ghosts);
      2. one synchronous Boid.step (ops.boids_step, csrc/continuous.cu) over owned + ghost boids -- ghosts keep
         their global coordinates, the kernel's torus minimum-image arithmetic needs no shift; only the owned
         results are kept (ghosts are ~2*vision/slab of the work);
      3. migration: boids whose new y left the slab move to the neighbouring rank with their global id.
    The exchanges are NCCL point-to-point messages with a counts exchange first (variable sizes).  The reference
    has no counterpart (single process, O(N^2) scan per step: continuous_space.py:202-235)."""

    def __init__(self, width: float, height: float, positions, directions, ids, *, vision=10.0, separation=2.0,
                 cohere=0.03, separate=0.015, match=0.05, speed=1.0, device=None, group=None):
        if not dist.is_initialized():
            raise RuntimeError("ShardedBoids needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.size = (float(width), float(height))
        self.params = dict(vision=float(vision), separation=float(separation), cohere=float(cohere),
                           separate=float(separate), match=float(match), speed=float(speed))
        self.slab = self.size[1] / self.world
        if self.world > 1 and self.slab < 2 * (vision + speed):
            raise ValueError("slabs must be at least 2*(vision+speed) high")
        self.y_lo, self.y_hi = self.rank * self.slab, (self.rank + 1) * self.slab
        self.pos = torch.as_tensor(positions, dtype=torch.float32).to(self.device).contiguous()
        self.dir = torch.as_tensor(directions, dtype=torch.float32).to(self.device).contiguous()
        self.ids = torch.as_tensor(ids, dtype=torch.int64).to(self.device).contiguous()
        self._ws = None
        self._ws_n = 0

    @staticmethod
    def owner_of(y: torch.Tensor, height: float, world: int) -> torch.Tensor:
        """Rank owning each y (slab index, the top edge y == height belongs to the last slab)."""
        return torch.clamp((y.to(torch.float64) * world / height).floor().to(torch.int64), 0, world - 1)

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])

    def _neighbours(self):
        return (self.rank - 1) % self.world, (self.rank + 1) % self.world

    def _exchange(self, to_down: torch.Tensor, to_up: torch.Tensor):
        """Send rows to the rank below / above, receive theirs; returns (from_down, from_up)."""
        down, up = self._neighbours()
        width = to_down.shape[1]
        counts = torch.zeros((self.world, 2), dtype=torch.int64, device=self.device)
        counts[self.rank, 0], counts[self.rank, 1] = to_down.shape[0], to_up.shape[0]
        dist.all_reduce(counts, group=self.group)
        counts = counts.tolist()
        n_from_down, n_from_up = counts[down][1], counts[up][0]   # what `down` sends up, what `up` sends down
        from_down = torch.empty((n_from_down, width), dtype=to_down.dtype, device=self.device)
        from_up = torch.empty((n_from_up, width), dtype=to_up.dtype, device=self.device)
        g = lambda r: dist.get_global_rank(self.group, r) if self.group is not dist.group.WORLD else r
        opsl = []
        # identical pair order on both ends of each link (matters when world == 2: one peer, two messages)
        first_up = self.rank % 2 == 0
        pairs = [("up", to_up, from_up, up), ("down", to_down, from_down, down)]
        if not first_up:
            pairs.reverse()
        for _, snd, rcv, peer in pairs:
            if snd.shape[0] > 0:
                opsl.append(dist.P2POp(dist.isend, snd.contiguous(), g(peer), self.group))
            if rcv.shape[0] > 0:
                opsl.append(dist.P2POp(dist.irecv, rcv, g(peer), self.group))
        if opsl:
            for q in dist.batch_isend_irecv(opsl):
                q.wait()
        return from_down, from_up

    def step(self) -> None:
        p = self.params
        reach = p["vision"]
        if self.world > 1:
            y = self.pos[:, 1]
            packed = torch.cat([self.pos, self.dir], dim=1)
            ghosts_down, ghosts_up = self._exchange(packed[y < self.y_lo + reach], packed[y >= self.y_hi - reach])
            allp = torch.cat([packed, ghosts_down, ghosts_up])
            pos_all, dir_all = allp[:, :2].contiguous(), allp[:, 2:].contiguous()
        else:
            pos_all, dir_all = self.pos, self.dir
        n_all = pos_all.shape[0]
        if self._ws is None or self._ws_n < n_all:
            self._ws_n = int(n_all * 1.2) + 1024
            self._ws = ops.CellListWorkspace(self._ws_n, self.size, p["vision"], self.device)
        new_pos, new_dir = ops.boids_step(pos_all, dir_all, self.size, torus=True, ws=self._ws, **p)
        n = self.n
        self.pos, self.dir = new_pos[:n], new_dir[:n]
        if self.world > 1:
            owner = self.owner_of(self.pos[:, 1], self.size[1], self.world)
            down, up = self._neighbours()
            stay = owner == self.rank
            go_down = (owner == down) & ~stay
            go_up = (owner == up) & ~stay & ~go_down
            if bool((~(stay | go_down | go_up)).any()):
                raise RuntimeError("a boid jumped over a whole slab (speed too large for this decomposition)")
            # ids travel exactly: pack them as two float32-sized int32 words
            idw = torch.stack([(self.ids & 0xFFFFFFFF).to(torch.int32), (self.ids >> 32).to(torch.int32)], dim=1).view(torch.float32)
            rec = torch.cat([self.pos, self.dir, idw], dim=1)
            from_down, from_up = self._exchange(rec[go_down], rec[go_up])
            rec = torch.cat([rec[stay], from_down, from_up])
            self.pos, self.dir = rec[:, :2].contiguous(), rec[:, 2:4].contiguous()
            w = rec[:, 4:6].contiguous().view(torch.int32).to(torch.int64)
            self.ids = (w[:, 0] & 0xFFFFFFFF) | (w[:, 1] << 32)
        else:
            self.pos, self.dir = self.pos.contiguous(), self.dir.contiguous()

    def gather(self):
        """(ids, pos, dir) of all boids on every rank, sorted by id (tests / readback)."""
        n = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        n[self.rank] = self.n
        dist.all_reduce(n, group=self.group)
        width = int(n.max().item())
        mine = torch.zeros((width, 5), dtype=torch.float64, device=self.device)
        mine[: self.n, :2], mine[: self.n, 2:4], mine[: self.n, 4] = self.pos, self.dir, self.ids
        everyone = torch.empty((self.world, width, 5), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        rows = torch.cat([everyone[r, : int(n[r].item())] for r in range(self.world)])
        order = ops.argsort_stable(rows[:, 4].contiguous())
        rows = rows[order]
        return rows[:, 4].to(torch.int64), rows[:, :2].to(torch.float32), rows[:, 2:4].to(torch.float32)



class ShardedBoids:
    """BoidFlockers on a continuous torus sharded by spatial domain (BASELINE.json config 4),
    with the exchanges done by kernels that write into the neighbours' memory over NVLink.

    Rank k owns the slab  y in [k, k+1) * height / world.  Per step:
      1. `mb_boids_send_ghosts`: boids within `vision` of a slab edge are appended to the neighbour's ghost inbox
         (peer-mapped symmetric memory, atomic slot counter in the receiver's memory); device barrier;
      2. one synchronous Boid.step (`ops.boids_step`) over owned + ghost boids (ghosts keep global coordinates:
         the torus minimum-image arithmetic needs no shift); only the owned results are kept;
      3. `mb_boids_route` / `mb_boids_compact`: boids whose new y left the slab are appended to their new owner's
         migrant inbox with their global id, the others are compacted in place (stable); device barrier; the
         migrants are appended.
    No NCCL call and no torch-level packing on the step path; the host reads two small counter vectors per step."""

    def __init__(self, width: float, height: float, positions, directions, ids, *, vision=10.0, separation=2.0,
                 cohere=0.03, separate=0.015, match=0.05, speed=1.0, device=None, group=None, capacity=None,
                 inbox=None):
        import torch.distributed._symmetric_memory as symm_mem

        if not dist.is_initialized():
            raise RuntimeError("ShardedBoids needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = dev = torch.device(device if device is not None else torch.cuda.current_device())
        self.size = (float(width), float(height))
        self.params = dict(vision=float(vision), separation=float(separation), cohere=float(cohere),
                           separate=float(separate), match=float(match), speed=float(speed))
        self.slab = self.size[1] / self.world
        if self.world > 1 and self.slab < 2 * (vision + speed):
            raise ValueError("slabs must be at least 2*(vision+speed) high")
        self.y_lo, self.y_hi = self.rank * self.slab, (self.rank + 1) * self.slab
        pos = torch.as_tensor(positions, dtype=torch.float32).to(dev)
        n = int(pos.shape[0])
        nmax = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=self.group)
        nmax = int(nmax.item())
        # inbox capacity: 4x the boids expected within `vision` of an edge (uniform density), at least 4096
        self.icap = int(inbox) if inbox is not None else max(4096, int(4 * nmax * vision / max(self.slab, vision)) + 1024)
        self.cap = int(capacity) if capacity is not None else int(nmax * 1.5) + 4 * self.icap + 1024
        self.n = n
        self.pos = [torch.empty((self.cap, 2), dtype=torch.float32, device=dev) for _ in range(2)]
        self.dir = [torch.empty((self.cap, 2), dtype=torch.float32, device=dev) for _ in range(2)]
        self.ids = [torch.empty(self.cap, dtype=torch.int64, device=dev) for _ in range(2)]
        self.pos[0][:n] = pos
        self.dir[0][:n] = torch.as_tensor(directions, dtype=torch.float32).to(dev)
        self.ids[0][:n] = torch.as_tensor(ids, dtype=torch.int64).to(dev)
        # symmetric inboxes: rows [0, icap) are written by the rank below me, rows [icap, 2*icap) by the rank above
        def symmetric(shape, dtype):
            t = symm_mem.empty(shape, dtype=dtype, device=dev)
            return t, symm_mem.rendezvous(t, self.group)

        self.gpos, self._gpos_h = symmetric((2 * self.icap, 2), torch.float32)
        self.gdir, self._gdir_h = symmetric((2 * self.icap, 2), torch.float32)
        self.mpos, self._mpos_h = symmetric((2 * self.icap, 2), torch.float32)
        self.mdir, self._mdir_h = symmetric((2 * self.icap, 2), torch.float32)
        self.mids, self._mids_h = symmetric((2 * self.icap,), torch.int64)
        self.cnt, self._cnt_h = symmetric((8,), torch.int32)   # [ghosts from below, ghosts from above, migrants ...]
        self.cnt.zero_()
        self.flags = torch.zeros(4, dtype=torch.int32, device=dev)
        self.stay = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.offs = torch.empty(self.cap, dtype=torch.int32, device=dev)
        words = ctypes.c_size_t(0)
        N.check(N.lib().mb_scan_scratch_words(self.cap, ctypes.byref(words)))
        self.scratch = torch.empty(int(words.value), dtype=torch.int32, device=dev)
        self._ws = ops.CellListWorkspace(self.cap, self.size, self.params["vision"], dev)
        down, up = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        row = self.icap * 2 * 4          # byte offset of the "from above" half of a [2*icap, 2] float32 inbox
        P = lambda h, r, off=0: h.buffer_ptrs[r] + off
        if self.world > 1:
            # what I send DOWN lands in the lower neighbour's "from above" half, what I send UP in the upper one's "from below" half
            self._g_down = (P(self._gpos_h, down, row), P(self._gdir_h, down, row), P(self._cnt_h, down, 4))
            self._g_up = (P(self._gpos_h, up), P(self._gdir_h, up), P(self._cnt_h, up, 0))
            self._m_down = (P(self._mpos_h, down, row), P(self._mdir_h, down, row), P(self._mids_h, down, self.icap * 8), P(self._cnt_h, down, 12))
            self._m_up = (P(self._mpos_h, up), P(self._mdir_h, up), P(self._mids_h, up), P(self._cnt_h, up, 8))
        self.cur = 0
        torch.cuda.current_stream(dev).synchronize()
        dist.barrier(group=self.group)

    owner_of = staticmethod(ShardedBoidsNCCL.owner_of)

    def _check_flags(self, flags):
        if flags[0]:
            raise RuntimeError("ShardedBoids: an inbox overflowed (raise `inbox=`)")
        if flags[1]:
            raise RuntimeError("a boid jumped over a whole slab (speed too large for this decomposition)")

    def step(self) -> None:
        p, lib, dev = self.params, N.lib(), self.device
        cur, nxt = self.cur, 1 - self.cur
        pos, dirs, ids, n = self.pos[cur], self.dir[cur], self.ids[cur], self.n
        st = N.stream_ptr(dev)
        with torch.cuda.device(dev):
            g = 0
            if self.world > 1:
                N.check(lib.mb_boids_send_ghosts(pos.data_ptr(), dirs.data_ptr(), n, self.y_lo, self.y_hi, p["vision"],
                                                 *self._g_down, *self._g_up, self.icap, self.flags.data_ptr(), st))
                self._cnt_h.barrier()
                c = self.cnt[:2].tolist()
                gd, gu = int(c[0]), int(c[1])
                if gd > self.icap or gu > self.icap:
                    raise RuntimeError("ShardedBoids: a ghost inbox overflowed (raise `inbox=`)")
                g = gd + gu
                pos[n:n + gd] = self.gpos[:gd]
                dirs[n:n + gd] = self.gdir[:gd]
                pos[n + gd:n + g] = self.gpos[self.icap:self.icap + gu]
                dirs[n + gd:n + g] = self.gdir[self.icap:self.icap + gu]
                self.cnt[:2].zero_()
            ops.boids_step(pos[:n + g], dirs[:n + g], self.size, torus=True, ws=self._ws,
                           pos_out=self.pos[nxt][:n + g], dir_out=self.dir[nxt][:n + g], **p)
            if self.world == 1:
                self.ids[nxt][:n] = ids[:n]
                self.cur = nxt
                return
            # route the owned results: stayers are compacted back into the `cur` buffers, leavers go to the inboxes
            N.check(lib.mb_boids_route(self.pos[nxt].data_ptr(), self.dir[nxt].data_ptr(), ids.data_ptr(), n, self.size[1],
                                       self.world, self.rank, self.stay.data_ptr(), self.offs.data_ptr(),
                                       self.scratch.data_ptr(), *self._m_down, *self._m_up, self.icap,
                                       self.flags.data_ptr(), st))
            N.check(lib.mb_boids_compact(self.pos[nxt].data_ptr(), self.dir[nxt].data_ptr(), ids.data_ptr(), n,
                                         self.stay.data_ptr(), self.offs.data_ptr(), pos.data_ptr(), dirs.data_ptr(),
                                         self.ids[nxt].data_ptr(), st))
            self._cnt_h.barrier()
            info = torch.cat([self.cnt[2:4], self.flags[:2], self.offs[n - 1:n] + self.stay[n - 1:n]]).tolist() if n > 0 \
                else torch.cat([self.cnt[2:4], self.flags[:2], torch.zeros(1, dtype=torch.int32, device=dev)]).tolist()
            md, mu, ns = int(info[0]), int(info[1]), int(info[4])
            self._check_flags(info[2:4])
            if md > self.icap or mu > self.icap or ns + md + mu > self.cap - 2 * self.icap:
                raise RuntimeError("ShardedBoids: capacity exceeded")
            new_ids = self.ids[nxt]
            pos[ns:ns + md], dirs[ns:ns + md], new_ids[ns:ns + md] = self.mpos[:md], self.mdir[:md], self.mids[:md]
            lo = self.icap
            pos[ns + md:ns + md + mu], dirs[ns + md:ns + md + mu] = self.mpos[lo:lo + mu], self.mdir[lo:lo + mu]
            new_ids[ns + md:ns + md + mu] = self.mids[lo:lo + mu]
            self.cnt[2:4].zero_()
            self.n = ns + md + mu
            # positions / directions stay in the `cur` buffers, ids moved to the other id buffer
            self.ids[cur], self.ids[nxt] = self.ids[nxt], self.ids[cur]

    def gather(self):
        """(ids, pos, dir) of all boids on every rank, sorted by id (tests / readback)."""
        n = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        n[self.rank] = self.n
        dist.all_reduce(n, group=self.group)
        width = int(n.max().item())
        mine = torch.zeros((width, 5), dtype=torch.float64, device=self.device)
        c = self.cur
        mine[: self.n, :2], mine[: self.n, 2:4], mine[: self.n, 4] = self.pos[c][: self.n], self.dir[c][: self.n], self.ids[c][: self.n]
        everyone = torch.empty((self.world, width, 5), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        rows = torch.cat([everyone[r, : int(n[r].item())] for r in range(self.world)])
        rows = rows[ops.argsort_stable(rows[:, 4].contiguous())]
        return rows[:, 4].to(torch.int64), rows[:, :2].to(torch.float32), rows[:, 2:4].to(torch.float32)
