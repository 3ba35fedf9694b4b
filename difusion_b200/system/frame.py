"""``FramePipeline`` - one frame of the SLAM loop (reference main.py:71-94: tracker linearisation against the map built so
far, then ``integrate_keyframe``) as ONE replayable CUDA graph.

The reference bakes every per-frame quantity into host state (tensor shapes, ``Isometry`` objects), so each frame is ~100
separately launched torch ops with ~25 host syncs.  Here the per-frame values live in a device block (``dif_frame_params``:
point count, frame number, the two poses) in front of the frame's points, so the launch sequence of ``dif_frame`` is
identical for every frame: it is captured once and replayed with one ``cudaGraphLaunch``.  Per frame the host does

    1 copy (header + points, contiguous, exactly 128 + 36 n bytes)  ->  1 graph launch  ->  [1 event sync when it reads]

and the graph's last node copies the 400-byte result block (H, g, energy, valid count, integrate counters) into pinned
host memory.  The map tensors must keep their addresses while a graph is alive: the pipeline re-captures when the map grew.

This class adds nothing to the reference's API surface; ``DenseIndexedMap.integrate_keyframe`` / ``SDFTracker.compute_sdf_Hg``
remain the drop-in calls.  It is the fast path for callers that own the loop (bench.py, tools/).
"""
from __future__ import annotations

import ctypes
import os
from collections import deque

import numpy as np
import torch

from .. import _lib

HDR = _lib.FRAME_HEADER_FLOATS
ROW = _lib.FRAME_POINT_FLOATS


def pack_frame(out: np.ndarray, seq: int, pc_cam: np.ndarray, xyz_world: np.ndarray, normal_world: np.ndarray, R_last, t_last,
               R_delta=None, t_delta=None) -> int:
    """Fill a host float32 buffer [HDR + ROW * n] with the frame header and the interleaved point rows; returns the float count."""
    n = int(pc_cam.shape[0])
    hdr = out[:HDR]
    hdr[:] = 0.0
    hi = hdr.view(np.int32)
    hi[0], hi[1] = n, seq
    hdr[4:13] = np.asarray(R_last, np.float32).ravel()
    hdr[13:16] = np.asarray(t_last, np.float32).ravel()
    hdr[16:25] = np.eye(3, dtype=np.float32).ravel() if R_delta is None else np.asarray(R_delta, np.float32).ravel()
    hdr[25:28] = 0.0 if t_delta is None else np.asarray(t_delta, np.float32).ravel()
    rows = out[HDR:HDR + ROW * n].reshape(n, ROW)
    rows[:, 0:3], rows[:, 3:6], rows[:, 6:9] = pc_cam, xyz_world, normal_world
    return HDR + ROW * n


class FramePipeline:
    N_BUF = 2                      # staging buffers: frame f+1 uploads while frame f computes
    MAX_AHEAD = 4                  # frames the host may run ahead of the device before it waits for the oldest

    def __init__(self, map, max_points: int, huber_k: float = 5.0, use_graph: bool = None):
        self.map, self.max_points, self.huber_k = map, int(max_points), float(huber_k)
        self.use_graph = (os.environ.get("DIF_FRAME_GRAPH", "1") != "0") if use_graph is None else bool(use_graph)
        dev = map.device
        self.device = dev
        L = map._L
        with torch.cuda.device(dev):
            self.stage = [torch.zeros(HDR + ROW * self.max_points, dtype=torch.float32, device=dev) for _ in range(self.N_BUF)]
            self.result = torch.zeros(_lib.FRAME_RESULT_BYTES, dtype=torch.uint8, device=dev)
            self.unq_mask = torch.zeros(self.max_points, dtype=torch.uint8, device=dev)
            self.scratch = torch.empty(L.dif_integrate_scratch_bytes(self.max_points), dtype=torch.uint8, device=dev)
            self.icp_scratch = torch.zeros(L.dif_icp_scratch_bytes(self.max_points), dtype=torch.uint8, device=dev)
            self.copy_stream = torch.cuda.Stream(device=dev)
        self.result_host = torch.zeros(_lib.FRAME_RESULT_BYTES, dtype=torch.uint8).pin_memory()
        self._res_np = self.result_host.numpy()
        self._copied = [torch.cuda.Event() for _ in range(self.N_BUF)]
        self._consumed = [torch.cuda.Event() for _ in range(self.N_BUF)]
        for e in self._consumed:
            e.record(torch.cuda.current_stream(dev))
        self._n_points_k = [0] * self.N_BUF
        self._ev_pool = [torch.cuda.Event() for _ in range(self.MAX_AHEAD + 2)]
        self._ev_next = 0
        self._inflight = deque()           # done-events of launched frames, oldest first
        self._graphs, self._graph_key = {}, None
        self._next = 0
        self._map_epoch = map._reset_epoch
        self.n_captures = 0

    # ------------------------------------------------------------------ the launch sequence (captured or direct)
    def _enqueue(self, k: int, flags: int):
        m = self.map
        view = m._view()
        st = _lib.stream_ptr(self.device)
        stage = self.stage[k]
        _lib.check(m._L.dif_frame(ctypes.byref(view), m._prep.encoder.data_ptr(), m._prep.decoder.data_ptr(),
                                  stage.data_ptr() + 4 * HDR, self.max_points, stage.data_ptr(), self.huber_k, flags,
                                  self.unq_mask.data_ptr(), m._persist.data_ptr(), m._persist.numel(), self.scratch.data_ptr(),
                                  self.scratch.numel(), self.icp_scratch.data_ptr(), self.icp_scratch.numel(), self.result.data_ptr(), st),
                   "dif_frame")
        self.result_host.copy_(self.result, non_blocking=True)

    def _graph_for(self, k: int, flags: int):
        m = self.map
        key = (m._latent.data_ptr(), m._persist.data_ptr(), m._cap_phys)
        if self._graph_key != key:
            self._graphs, self._graph_key = {}, key
        g = self._graphs.get((k, flags))
        if g is not None:
            return g
        torch.cuda.synchronize(self.device)                 # (every stream: an upload into this staging buffer may be in flight)
        # warm-up outside capture (module load, function attributes must not happen while capturing): a frame with zero points
        # leaves the map untouched.  The staging buffer may already hold the uploaded frame: save and restore its header.
        hdr = torch.zeros(HDR, dtype=torch.float32)
        hdr[4], hdr[8], hdr[12], hdr[16], hdr[20], hdr[24] = 1, 1, 1, 1, 1, 1
        saved = self.stage[k][:HDR].clone()
        self.stage[k][:HDR].copy_(hdr.to(self.device))
        self._enqueue(k, flags)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue(k, flags)
        self.stage[k][:HDR].copy_(saved)
        torch.cuda.synchronize(self.device)
        self._graphs[(k, flags)] = g
        self.n_captures += 1
        return g

    # ------------------------------------------------------------------ per-frame calls
    def _reserve(self, n_points: int):
        """Capacity rule of integrate_keyframe: a frame allocates at most 7 PLIVoxes per point.  The host never runs more than
        MAX_AHEAD frames ahead of the device (it waits for the OLDEST frame in flight, so the device queue never drains), and the
        frames whose counters it has not read yet are accounted with the same bound; the map therefore grows (and the graphs are
        re-captured) only at a point where the device is idle anyway."""
        m = self.map
        while len(self._inflight) > self.MAX_AHEAD:
            self._inflight.popleft().synchronize()
        if self._inflight and self._inflight[0].query():
            while self._inflight and self._inflight[0].query():
                self._inflight.popleft()
        self._peek_counters()
        worst = min(7 * n_points, m._n_cells)
        if m._n_occ_host + (len(self._inflight) + 1) * worst > m._cap_phys:
            self.sync()
            if m._n_occ_host + worst > m._cap_phys:
                from .map import _next_pow2
                m._grow(_next_pow2(m._n_occ_host + (self.MAX_AHEAD + 2) * worst))

    def _peek_counters(self):
        """n_occupied only ever grows: the value in the pinned result block (written by whichever frame finished last) is a valid
        lower bound of the current one without any synchronisation."""
        if self._map_epoch != self.map._reset_epoch:      # the map was reset in place: the block still holds the old map's counters
            self._map_epoch = self.map._reset_epoch
            self._res_np[352:] = 0
        n = int(self._res_np[352:].view(np.int32)[_lib.STAT_N_OCCUPIED])
        if n > self.map._n_occ_host:
            self.map._n_occ_host = n

    def upload(self, packed: torch.Tensor, n_points: int) -> int:
        """Copy a packed frame ([header | rows], see pack_frame; pinned host or device memory) into the next staging buffer on the
        copy stream; returns the buffer index to pass to launch().  Call order per frame: launch(f), upload(f+1), sync(f) - the copy of
        frame f+1 then overlaps frame f's kernels AND its host-side cost (~12 us of stream / event calls) is off the launch path
        (issuing the upload before the launch cost 28 us per frame end to end: 6 416 -> 7 836 frames/s in bench.py)."""
        assert n_points <= self.max_points and packed.numel() == HDR + ROW * n_points
        k = self._next
        self._next = (k + 1) % self.N_BUF
        self.copy_stream.wait_event(self._consumed[k])               # the previous frame that used this buffer has run
        with torch.cuda.stream(self.copy_stream):
            self.stage[k][:packed.numel()].copy_(packed, non_blocking=True)
            self._copied[k].record(self.copy_stream)
        self._n_points_k[k] = n_points
        return k

    def launch(self, k: int, track: bool = True, integrate: bool = True):
        """Run the frame uploaded into buffer k on the current stream (graph replay, or the direct launch sequence)."""
        flags = (_lib.FRAME_TRACK if track else 0) | (_lib.FRAME_INTEGRATE if integrate else 0)
        main = torch.cuda.current_stream(self.device)
        with self.map.modifying_lock:
            if integrate:
                self._reserve(self._n_points_k[k])
            g = self._graph_for(k, flags) if self.use_graph else None
            main.wait_event(self._copied[k])
            if g is not None:
                g.replay()
            else:
                self._enqueue(k, flags)
            self._consumed[k].record(main)
            ev = self._ev_pool[self._ev_next]
            self._ev_next = (self._ev_next + 1) % len(self._ev_pool)
            ev.record(main)
            self._inflight.append(ev)
            self._last_flags = flags

    def sync(self):
        """Wait for the last launched frame; returns (icp (44,) float64 view, counters list) of that frame and feeds the map's counters."""
        launched = bool(self._inflight)
        while self._inflight:
            ev = self._inflight.pop()                    # the newest event covers all older ones (one stream)
            ev.synchronize()
            self._inflight.clear()
        icp = self._res_np[:352].view(np.float64)
        st = self._res_np[352:].view(np.int32).tolist()
        if launched and (getattr(self, "_last_flags", 0) & _lib.FRAME_INTEGRATE):
            self.map._consume_stats(st)
        if self.map._pending_error is not None:
            err, self.map._pending_error = self.map._pending_error, None
            raise err
        return icp, st
