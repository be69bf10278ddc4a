"""Sync-free frame engine for the SST hot path (BASELINE config 2):

    points --voxelize--> DynamicVFE --> SSTInputLayerV2 (2 window plans) --> SSTv2 (num_blocks x 2 SRA layers)

It drives the same libsstb200 entry points as the registered modules, but with device-resident row counts
(`n_dev`) and capacity-sized buffers, so the whole frame is one CUDA graph: no host sync, no allocation, one
launch per frame.  Mirrors DynamicVoxelNet.extract_feat (mmdet3d/models/detectors/dynamic_voxelnet.py:38-47)
up to and including the backbone's sparse output.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops
from .sst_modules import PRECISIONS, _SraPlan, SSTInputLayerV2, SSTv2
from .voxel_modules import DynamicVFE

L.SIGNATURES["sstb200_branch_fork"] = (C.c_int, [L.vp, L.vp])
L.SIGNATURES["sstb200_branch_join"] = (C.c_int, [L.vp, L.vp])
L.SIGNATURES["sstb200_graph_begin"] = (C.c_int, [L.vp])
L.SIGNATURES["sstb200_graph_end"] = (C.c_int, [L.vp, C.POINTER(C.c_void_p), L.P_i32, L.P_i32])
L.SIGNATURES["sstb200_graph_launch"] = (C.c_int, [L.vp, L.vp])
L.SIGNATURES["sstb200_graph_destroy"] = (C.c_int, [L.vp, L.vp])
L.SIGNATURES["sstb200_voxelize_frames"] = (C.c_int, [L.vp, L.vp, C.c_int, C.c_int, L.vp, C.c_int, L.P_f32, L.P_f32, L.vp])


class SSTEngine:
    def __init__(self, voxel_size, point_cloud_range, voxel_encoder: DynamicVFE, middle_encoder: SSTInputLayerV2,
                 backbone: SSTv2, max_points, batch_size=1, precision="fp32", device="cuda:0", use_graph=True):
        self.dev = torch.device(device)
        self.vs = [float(v) for v in voxel_size]
        self.rng = [float(v) for v in point_cloud_range]
        self.vfe, self.il, self.bb = voxel_encoder.eval(), middle_encoder.eval(), backbone.eval()
        self.cap, self.B = int(max_points), int(batch_size)
        self.precision = precision
        self.F = self.vfe.raw_in_channels
        self.d = backbone.d_model[0]
        assert all(d == self.d for d in backbone.d_model), "engine assumes a constant d_model"
        assert not hasattr(backbone, "linear0"), "linear0 not wired into the engine"
        if voxel_encoder.feat_channels[-1] != self.d:
            raise ValueError(f"voxel encoder emits {voxel_encoder.feat_channels[-1]} channels but the backbone expects d_model = {self.d} "
                             "(no linear0 in the engine path)")
        self.il.set_drop_info()
        if self.il._may_drop():
            raise NotImplementedError("the sync-free engine covers drop_info settings that cannot drop voxels "
                                      "(eval configs); use the modules for training-time dropping")
        dev, cap = self.dev, self.cap
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        self.points = torch.zeros((cap, self.F), **f32)
        self.offsets = torch.zeros((self.B + 1,), **i32)
        self.coors4 = torch.empty((cap, 4), **i32)
        self.vf = torch.empty((cap, self.vfe.feat_channels[-1]), **f32)
        self.vc = torch.empty((cap, 4), **i32)
        self.num = torch.zeros((1,), **i32)
        self.x = [torch.empty((cap, self.d), **f32) for _ in range(2)]
        self.plans = []
        for _ in range(2):
            p = {k: torch.empty((cap + 16,) if k in ("win_offsets", "win_batch") else (cap,), **i32)
                 for k in ("pos_code", "tok_win", "tok_inner", "win_offsets", "tok_perm", "win_level", "win_rank", "tok_slot",
                           "win_batch")}
            p["counters"] = torch.zeros((20,), **i32)
            self.plans.append(p)
        self.wcfg, _ = ops._window_cfg(self.il.sparse_shape, self.il.window_shape, self.il.drop_info, self.B)
        tab, ndim, maxw, Lp = self.il._pos(self.d, dev)
        self.pos = (tab, ndim, maxw, Lp)
        self._shift_structs = [ops._WindowShift(None, None, None, None, p["pos_code"].data_ptr(), p["tok_win"].data_ptr(),
                                                p["tok_inner"].data_ptr(), p["win_offsets"].data_ptr(),
                                                p["tok_perm"].data_ptr(), p["win_level"].data_ptr(),
                                                p["win_rank"].data_ptr(), p["counters"].data_ptr(), p["tok_slot"].data_ptr(),
                                                p["win_batch"].data_ptr())
                               for p in self.plans]
        self._plan_structs = [_SraPlan(p["win_offsets"].data_ptr(), p["tok_perm"].data_ptr(), p["tok_win"].data_ptr(),
                                       p["pos_code"].data_ptr(), p["counters"].data_ptr(), tab.data_ptr(), Lp, maxw, ndim,
                                       max(v["max_tokens"] for v in self.il.drop_info.values()), p["tok_slot"].data_ptr(),
                                       p["win_batch"].data_ptr())
                              for p in self.plans]
        self.vfe_cfg = self.vfe._cfg(self.B)
        self.vfe_cfg.precision = PRECISIONS[precision]
        self._build_layer_structs()
        self._vs_arr, self._rng_arr = L.arr(C.c_float, self.vs), L.arr(C.c_float, self.rng)
        self.stream = torch.cuda.Stream(device=dev)
        self.side = torch.cuda.Stream(device=dev)   # side branch: the two window plans run next to the VFE layers
        self.graph = None
        self.launches_per_frame = None
        with torch.cuda.stream(self.stream):
            self._enqueue()           # warm-up: grows the arena, sets kernel attributes
            self._enqueue()
        self.stream.synchronize()
        if use_graph:
            with torch.cuda.stream(self.stream):
                lib, c = L.lib(), L.ctx(self.dev)
                L.check(c, lib.sstb200_graph_begin(c))
                try:
                    self._enqueue()
                finally:
                    ex, nk, no = C.c_void_p(), C.c_int32(0), C.c_int32(0)
                    L.check(c, lib.sstb200_graph_end(c, C.byref(ex), C.byref(nk), C.byref(no)))
                self.graph = ex
                self.launches_per_frame = nk.value   # kernel nodes of OUR library in one frame
                self.other_nodes_per_frame = no.value
            self.stream.synchronize()

    def _build_layer_structs(self):
        """Layer descriptors for the C ABI.  The engine OWNS the fp16 weight copies whose addresses go into the descriptors (and
        from there into the captured graph): they stay alive as long as the engine does, whatever the modules do later."""
        from .sst_modules import _SraLayer
        prec = PRECISIONS[self.precision]
        layers = [layer for blk in self.bb.block_list for layer in blk.encoder_list]
        assert all(l.d_model == self.d for l in layers), "engine assumes a constant d_model"
        self._half_refs = [l.half_weights(refresh=True) for l in layers] if prec == 1 else []
        self._layers = [(layer._struct(prec), i) for blk in self.bb.block_list for i, layer in enumerate(blk.encoder_list)]
        self._layer_array = (_SraLayer * max(len(self._layers), 1))(*[ls for ls, _ in self._layers])

    def refresh_weights(self):
        """Call after the modules' parameters changed (optimizer step, load_state_dict): fp32 parameters are read in place by
        the graph, the fp16 copies of the tensor-core path are re-cast INTO the buffers the graph already points at."""
        if PRECISIONS[self.precision] != 1:
            return
        layers = [layer for blk in self.bb.block_list for layer in blk.encoder_list]
        with torch.no_grad():
            for l, refs in zip(layers, self._half_refs):
                sa = l.win_attn.self_attn
                for dst, src in zip(refs, (sa.in_proj_weight, sa.out_proj.weight, l.linear1.weight, l.linear2.weight)):
                    dst.copy_(src)

    # everything below is stream-ordered and sync-free
    def _enqueue(self):
        lib, c = L.lib(), L.ctx(self.dev)
        prec = PRECISIONS[self.precision]
        cap = self.cap
        L.check(c, lib.sstb200_voxelize_frames(c, self.points.data_ptr(), cap, self.F, self.offsets.data_ptr(), self.B,
                                               self._vs_arr, self._rng_arr, self.coors4.data_ptr()))
        L.check(c, lib.sstb200_dynamic_vfe_forward(c, C.byref(self.vfe_cfg), self.points.data_ptr(), self.coors4.data_ptr(),
                                                   cap, self.vf.data_ptr(), self.vc.data_ptr(), self.num.data_ptr(), None))
        # fork: the window plans need only the voxel coordinates (final after the first third of the VFE call)
        with torch.cuda.stream(self.side):
            cs = L.ctx(self.dev)
        L.check(c, lib.sstb200_branch_fork(c, cs))
        for s in range(2):
            L.check(cs, lib.sstb200_window_plan_i32(cs, self.vc.data_ptr(), cap, self.num.data_ptr(), C.byref(self.wcfg), s,
                                                    C.byref(self._shift_structs[s])))
        L.check(c, lib.sstb200_branch_join(c, cs))
        if self._layers:
            L.check(c, lib.sstb200_sra_stack_forward(c, self._layer_array, len(self._layers), C.byref(self._plan_structs[0]),
                                                     C.byref(self._plan_structs[1]), self.vf.data_ptr(), self.x[0].data_ptr(),
                                                     self.x[1].data_ptr(), cap, self.num.data_ptr(), prec))
            self._last = self.x[0]
        else:
            self._last = self.vf

    def load_frames_device(self, points_dev, offsets_dev):
        """points_dev [P,F] fp32 (frames back to back) and offsets_dev int32 [B+1], both already on the device.
        Stream-ordered D2D copy into the engine's input buffer (no sync)."""
        n = points_dev.shape[0]
        assert n <= self.cap and offsets_dev.numel() == self.B + 1
        # the caller produced points_dev / offsets_dev on ITS current stream: order the copies after that work
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            self.points[:n].copy_(points_dev, non_blocking=True)
            self.offsets.copy_(offsets_dev, non_blocking=True)

    def status(self):
        """Synchronises the engine stream and raises if the last frame hit a condition the module path would have raised on:
        a voxel outside the window grid (window-plan status word, one per shift).  The sync-free path itself never reads it."""
        self.stream.synchronize()
        for s, p in enumerate(self.plans):
            flag = int(p["counters"][18].item())
            if flag & 1:
                raise L.SSTB200Error(f"shift {s}: a voxel coordinate lies outside the window grid (sparse_shape too small for the input)")
        return True

    def run(self):
        """Enqueue one forward over the resident frames (graph replay).  Returns (feats_buf, coors_buf, num_dev) - PERSISTENT
        buffers of the engine: they are valid until the next run() on this engine; a consumer on another stream must order
        itself after `engine.stream` (e.g. `torch.cuda.current_stream().wait_stream(engine.stream)`) and finish before then."""
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                c = L.ctx(self.dev)
                L.check(c, L.lib().sstb200_graph_launch(c, self.graph))
            else:
                self._enqueue()
        return self._last, self.vc, self.num

    def forward_host(self, pinned_points, offsets_pinned, out_feats_pinned, out_coors_pinned):
        """End-to-end call on HOST buffers: H2D(points) -> forward -> D2H(num) -> D2H(feats[:M], coors[:M]).
        Returns M; the output copies are stream-ordered (synchronise the engine stream before reading them)."""
        self.submit_host(pinned_points, offsets_pinned, out_feats_pinned, out_coors_pinned)
        return self.collect_host(out_feats_pinned, out_coors_pinned)

    def submit_host(self, pinned_points, offsets_pinned, out_feats_pinned=None, out_coors_pinned=None):
        """Asynchronous half of forward_host: enqueue H2D + forward + D2H of the row count; returns immediately, so
        several engines (streams) can be kept in flight by one host thread.

        With the pinned output buffers given here, the D2H of the result is enqueued right behind the forward for a predicted
        row count (the largest of the recent frames + 2 %): the data-dependent size then costs no host round trip on the
        stream's critical path; `collect_host` copies the few missing rows in the rare case the prediction was short."""
        n = pinned_points.shape[0]
        if not hasattr(self, "_num_pinned"):
            self._num_pinned = torch.zeros((1,), dtype=torch.int32).pin_memory()
            self._done = torch.cuda.Event()
            self._rows_guess = 0
        self._rows_sent = 0
        with torch.cuda.stream(self.stream):
            self.points[:n].copy_(pinned_points, non_blocking=True)
            self.offsets.copy_(offsets_pinned, non_blocking=True)
            self.run()
            self._num_pinned.copy_(self.num, non_blocking=True)
            if out_feats_pinned is not None and self._rows_guess > 0:
                g = min(self._rows_guess, self.cap, out_feats_pinned.shape[0])
                out_feats_pinned[:g].copy_(self._last[:g], non_blocking=True)
                out_coors_pinned[:g].copy_(self.vc[:g], non_blocking=True)
                self._rows_sent = g
            self._done.record(self.stream)

    def collect_host(self, out_feats_pinned, out_coors_pinned):
        """Rows [0, M) of the pinned buffers receive the frame's voxel features / coordinates (rows beyond M are unspecified).
        They are complete on return when the prediction made at submit time covered M; otherwise the missing rows are copied
        stream-ordered (synchronise the engine stream before reading).  `d2h_rows` = rows that cross PCIe for this frame."""
        self._done.synchronize()            # the row count is data dependent: one small wait per frame
        M = int(self._num_pinned[0])
        sent = self._rows_sent
        if M > sent:                        # no prediction yet, or it was short: fetch the remainder
            with torch.cuda.stream(self.stream):
                out_feats_pinned[sent:M].copy_(self._last[sent:M], non_blocking=True)
                out_coors_pinned[sent:M].copy_(self.vc[sent:M], non_blocking=True)
        self.d2h_rows = max(M, sent)
        self._rows_guess = max(int(M * 1.02) + 64, int(self._rows_guess * 0.98))
        return M
