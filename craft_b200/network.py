"""CRAFT model with the reference's constructor, parameter names and forward() contract
(core/network.py:26-267), running its dense per-pixel path on the sm_100a kernels.

    CRAFT(args).forward(image1, image2, iters=12, flow_init=None, upsample=True, test_mode=0)

Drop-in notes: `args` is the same argparse.Namespace the reference drivers
build, and it is mutated the same way (corr_levels, corr_multiplier, *_trans_config, dropout);
`network.RAFTER` is the alias evaluate.py installs for old checkpoints.  fnet / cnet stay stock
PyTorch (outside the named hot path).  Forward-only in this round.
"""
import os

import torch
import torch.nn as nn

from . import hotpath as hp
from . import ops
from .corr import CorrBlock, TransCorrBlock
from .extractor import BasicEncoder
from .gma import Attention
from .ops import TokenGrid
import collections
import contextlib

from .setrans import SETransConfig, SelfAttVisPosTrans, WorkspaceCache, _require_inference
from .update import GMAUpdateBlock
from .utils.utils import coords_grid, print0, upflow8


class CRAFT(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.hidden_dim = hdim = 128
        self.context_dim = cdim = 128
        args.corr_levels = 4
        if "dropout" not in self.args:
            self.args.dropout = 0
        if args.corr_radius == -1:
            args.corr_radius = 4

        if args.craft:
            c = self.inter_trans_config = SETransConfig()
            c.update_config(args)
            c.in_feat_dim = c.feat_dim = 256
            c.max_pos_size = 160
            c.out_attn_scores_only = True
            c.attn_diag_cycles = 1000
            c.num_modes = args.inter_num_modes
            c.tie_qk_scheme = "shared"
            c.qk_have_bias = args.inter_qk_have_bias
            c.pos_code_type = args.inter_pos_code_type
            c.pos_code_weight = args.inter_pos_code_weight
            self.args.inter_trans_config = c
            self.corr_fn = TransCorrBlock(c, radius=self.args.corr_radius, do_corr_global_norm=True)

        self.fnet = BasicEncoder(output_dim=256, norm_fn="instance", dropout=args.dropout)
        self.cnet = BasicEncoder(output_dim=hdim + cdim, norm_fn="batch", dropout=args.dropout)

        if args.f2trans != "none":
            c = self.f2_trans_config = SETransConfig()
            c.update_config(args)
            c.in_feat_dim = c.feat_dim = 256
            c.has_input_skip = True
            c.has_FFN = False
            c.attn_mask_radius = args.f2_attn_mask_radius
            c.tie_qk_scheme = None
            c.qk_have_bias = False
            c.out_attn_probs_only = False
            c.attn_diag_cycles = 1000
            c.num_modes = args.f2_num_modes
            c.pos_code_type = args.intra_pos_code_type
            c.pos_code_weight = args.f2_pos_code_weight
            self.f2_trans = SelfAttVisPosTrans(c, "F2 transformer")
            self.args.f2_trans_config = c
            if args.f1trans != "none":
                raise NotImplementedError("--f1 shared/private (two-way correlation) is an ablation outside the "
                                          "hot path; the shipped checkpoints use f1trans='none'")
            self.f1_trans = None
            args.corr_multiplier = 1
        else:
            # network.py leaves corr_multiplier unset here; BasicMotionEncoder needs it
            if not hasattr(args, "corr_multiplier"):
                args.corr_multiplier = 1

        if args.use_setrans:
            c = self.intra_trans_config = SETransConfig()
            c.update_config(args)
            c.in_feat_dim = c.feat_dim = 128
            c.has_FFN = False
            c.has_input_skip = True
            c.attn_mask_radius = -1
            c.tie_qk_scheme = None
            c.qk_have_bias = False
            c.out_attn_probs_only = True
            c.attn_diag_cycles = 1000
            c.num_modes = args.intra_num_modes
            c.pos_code_type = args.intra_pos_code_type
            c.pos_code_weight = args.intra_pos_code_weight
            self.att = SelfAttVisPosTrans(c, "Intra-frame attention")
            self.args.intra_trans_config = c
        else:
            self.att = Attention(args=self.args, dim=cdim, heads=self.args.num_heads, max_pos_size=160, dim_head=cdim)

        self.update_block = GMAUpdateBlock(self.args, hidden_dim=hdim)
        self.call_counter = 0
        self.materialize_level0 = False    # True only for debugging / SAVECORR: writes the 205 MB fp32 level-0 volume
        self.encoder_nhwc = os.environ.get("CRAFT_B200_ENCODER_NHWC", "1") != "0"   # fused encoders -> token packer directly
        self.level0 = None                 # None: $CRAFT_B200_LEVEL0 or "h16"; "ondemand": level 0 is never stored
        # Inference calls are captured into one CUDA graph per (shape, iters, test_mode) and replayed:
        # the ~250 launches of a forward then cost no host time.  Set False (or CRAFT_B200_NO_GRAPH=1)
        # to launch eagerly.
        self.use_cuda_graph = os.environ.get("CRAFT_B200_NO_GRAPH", "0") != "1"
        # fnet / cnet are outside the hot path; TF32 convolutions there cost ~6e-3 max abs error on
        # features of magnitude 20 and are 2.5x faster than strict fp32 (profiles/README.md).
        # "fp16": encoder activations and conv operands in half precision (fp32 accumulation and fp32 norm
        # statistics) -- an 11-bit mantissa, i.e. finer than TF32's 10 bits, at half the HBM traffic.
        # The reference's own evaluation default is fp16 autocast (evaluate.py:1455-1456).
        # (both flags are set by set_precision below)
        # test_mode=1 returns only the LAST iteration's upsampled flow (core/network.py:262-263), yet the
        # reference computes the mask head and the convex upsampling in every iteration and drops them.
        # Eliding that dead work leaves every returned value bit-identical.
        # Set False to execute the reference's schedule literally.
        self.elide_dead_upsample = True
        # Precision tier -- the type of the tensor-core operands / activation buffers;
        # accumulators, statistics, GRU state, coordinates, flow and masks are fp32 in all of them:
        #   "fp16" (default)  float16 operands (11-bit mantissa), fp16 encoders.  Same speed as bf16 (measured:
        #                     205.1 vs 205.7 pairs/s) and 7x closer to the fp32 reference: mean EPE 1.6e-4 px at
        #                     448x1024, <= 3.9e-4 px on every trained-weight case -- inside north_star's fp32
        #                     tolerance (1e-3).  The reference's own evaluation default is fp16 autocast
        #                     (evaluate.py:1455-1456).
        #   "bf16"            bfloat16 operands (8-bit mantissa): north_star's 1e-2 px tier (1.1e-3 px at 448x1024);
        #                     wider exponent range, for inputs whose activations could exceed fp16's 65504.
        #   "fp32-parity"     float16 operands + strict-fp32 cuDNN encoders: 1.35e-4 px, 2.6x slower (encoders).
        self.set_precision(os.environ.get("CRAFT_B200_PRECISION", "fp16"))
        self._graphs = collections.OrderedDict()   # LRU, bounded by max_cached_graphs
        self.max_cached_graphs = 12
        self._workspaces = WorkspaceCache(capacity=12)  # this model's own device buffers
        self._lane = 0                                  # see on_lane()

    def set_precision(self, tier):
        if tier not in ("bf16", "fp16", "fp32-parity"):
            raise ValueError("precision tier must be 'bf16', 'fp16' or 'fp32-parity', got %r" % (tier,))
        self.precision = tier
        self.act_dtype = torch.bfloat16 if tier == "bf16" else torch.float16
        if tier == "fp32-parity":
            self.encoder_half, self.encoder_tf32 = False, False
        else:
            self.encoder_half = os.environ.get("CRAFT_B200_ENCODER", "fp16") == "fp16"
            self.encoder_tf32 = True
        return self

    @contextlib.contextmanager
    def on_lane(self, k):
        """Several independent pairs in flight on one GPU: inside `with model.on_lane(k):` a forward uses lane k's own
        workspace, CUDA graph and static input/output buffers (the parameters are shared), so forwards issued on
        different CUDA streams under different lanes may overlap on the device.  The refinement loop is a serial chain
        of kernels that each leave SMs idle (114 of 148 CTAs, prologue / epilogue bubbles); a second and third pair
        fill them: 242 -> 274 -> 284 pairs/s at 448x1024 (profiles/r02_lanes.txt).  craft_b200.pipeline.PairStream
        drives this; a plain `model(...)` call is lane 0.  (The attention diagnostics max_attn / clamp_count are
        shared by the lanes and updated without atomics.)"""
        prev, self._lane = self._lane, int(k)
        try:
            yield self
        finally:
            self._lane = prev

    def _ws_slot(self):
        return ("lane", self._lane) if self._lane else 0

    def workspace_for(self, H, W, device=None):
        """The device buffers this model uses for H x W images (token grid H/8 x W/8) at its precision tier --
        what tests and profiling scripts inspect after a forward."""
        device = device or next(self.parameters()).device
        with torch.cuda.device(device), ops.precision(self.act_dtype):
            return self._workspaces.get(TokenGrid(H // 8, W // 8), device, self.materialize_level0 or self.level0,
                                        slot=self._ws_slot())

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def initialize_flow(self, img):
        N, _, H, W = img.shape
        c = coords_grid(N, H // 8, W // 8, device=img.device)
        return c, c.clone()

    def upsample_flow(self, flow, mask):
        """[N,2,h,w], [N,576,h,w] -> [N,2,8h,8w] (core/network.py:151-162), via the upsample kernel."""
        N, _, h, w = flow.shape
        g = TokenGrid(h, w)
        with torch.cuda.device(flow.device):
            out = torch.empty((N, 2, 8 * h, 8 * w), dtype=torch.float32, device=flow.device)
            fr = torch.zeros((g.Mp, 2), dtype=torch.float32, device=flow.device)
            mr = torch.zeros((g.Mp, 576), dtype=torch.float32, device=flow.device)
            for b in range(N):
                fr.view(g.H, g.Wp, 2)[:, :g.W] = flow[b].float().permute(1, 2, 0)
                mr.view(g.H, g.Wp, 576)[:, :g.W] = mask[b].float().permute(1, 2, 0)
                ops.upsample_flow(mr, fr, g, out=out[b])
        return out

    # ------------------------------------------------------------------------------------------
    def _encoders(self, image1, image2):
        amp = bool(getattr(self.args, "mixed_precision", False))
        self.fnet.fused_half = self.cnet.fused_half = bool(self.encoder_half)
        if self.encoder_nhwc and not amp and self.fnet._can_fuse(image1) and self.cnet._can_fuse(image1):
            # nn.DataParallel replicas already run side by side (one thread per device): one stream each
            return self._encoders_fused(image1, image2, one_stream=getattr(self, "_is_replica", False))
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        # fnet (two frames) and cnet (frame 1) are independent: run cnet on a side stream so their
        # many small, latency-bound kernels overlap (fork/join is captured into the CUDA graph too).
        main = torch.cuda.current_stream()
        side = self._side_stream(image1.device)
        side.wait_stream(main)
        with torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark,
                                        deterministic=False, allow_tf32=self.encoder_tf32), \
                torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            with torch.cuda.stream(side):
                cnet_feat = self.cnet(image1).float().contiguous()
            fmap1, fmap2 = self.fnet([image1, image2])
        main.wait_stream(side)
        cnet_feat.record_stream(main)
        return fmap1.float().contiguous(), fmap2.float().contiguous(), cnet_feat

    def _encoders_fused(self, image1, image2, one_stream=False):
        """Inference fast path of both encoders: ONE input kernel (normalisation, 2x2
        space-to-depth, channels-last, the first convolution's zero border) for all three encoder inputs; cuDNN
        convolutions on channels-last tensors (the 7x7/2 one as a 4x4/1 over 16 channels); craft_b200 norm / ReLU /
        residual kernels; the channels-last output of the last convolution goes straight to the token packers."""
        B = image1.shape[0]
        s = ops.image_s2d(torch.cat([image1, image2], dim=0), dtype=self.fnet.fused_dtype())
        main = torch.cuda.current_stream()
        side = main if one_stream else self._side_stream(image1.device)
        side.wait_stream(main)
        with torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark,
                                        deterministic=False, allow_tf32=self.encoder_tf32):
            with torch.cuda.stream(side):      # fnet (two frames) and cnet (frame 1) are independent
                cn = self.cnet.forward_nhwc(s2d=s[:B])
            fm = self.fnet.forward_nhwc(s2d=s)
        main.wait_stream(side)
        cn.record_stream(main)
        return ([ops.NhwcFeat(fm[b]) for b in range(B)], [ops.NhwcFeat(fm[B + b]) for b in range(B)],
                [ops.NhwcFeat(cn[b]) for b in range(B)])

    def _side_stream(self, device):
        key = str(device)
        if not hasattr(self, "_streams"):
            self._streams = {}
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=device)
        return self._streams[key]

    def _prepare_pair(self, ws, fmap1, fmap2, cnet_feat, flow_init):
        """Everything that happens once per pair (core/network.py:179-228) on one batch element."""
        g = ws.grid
        a = self.args
        if a.f2trans != "none":
            att2 = self.f2_trans.attend(ws, fmap2, ws.T2, ws.Q2, ws.K2, ws.lse2_f2, ws.clip_f2, slot=1)
            self.f2_trans.setrans.out_trans.run(ws, att2, ws.T2, 0, out_b=ws.T2f)
            second_is_tokens = True
        else:
            second_is_tokens = False
        if a.craft:
            ops.pack_tokens(fmap1, g, ops.PACK_LN, out_b=ws.T1)
            if not second_is_tokens:
                ops.pack_tokens(fmap2, g, ops.PACK_LN, out_b=ws.T2f)
            # f2_trans already ends in the same affine-free LayerNorm the correlation encoder applies
            # (core/setrans.py:407 then :794): LN(LN(x)) == LN(x) up to eps = 1e-12.
            self.corr_fn.build_rows(ws, ws.T1, ws.T2f)
        else:
            ops.pack_tokens(fmap1, g, ops.PACK_COPY, out_b=ws.Qc)
            if second_is_tokens:
                ws.Kc.copy_(ws.T2f)
            else:
                ops.pack_tokens(fmap2, g, ops.PACK_COPY, out_b=ws.Kc)
            hp.build_correlation(ws, ws.Qc, ws.Kc, M=1, d=256, w_agg=0.0, table=None, w_pos=0.0, global_norm=False)
        if isinstance(cnet_feat, ops.NhwcFeat):
            net, inp = cnet_feat.window(0, 128), cnet_feat.window(128, 128)
        else:
            net, inp = cnet_feat[:128].contiguous(), cnet_feat[128:].contiguous()
        ops.pack_tokens(net, g, ops.PACK_TANH, out_b=ws.X, colb=0, out_f=ws.Hm)
        ops.pack_tokens(inp, g, ops.PACK_RELU, out_b=ws.X, colb=128)
        if a.use_setrans:
            att = self.att.attend(ws, inp, ws.Ta, ws.Qa, ws.Ka, ws.lse2_att, ws.clip_att, slot=2,
                                  pack_mode=ops.PACK_RELU_LN)
        else:
            ops.pack_tokens(inp, g, ops.PACK_RELU, out_b=ws.Ta)
            att = self.att.attend(ws, ws.Ta, ws.Qa, ws.Ka, ws.lse2_att, ws.clip_att)
        ops.init_coords(ws.coords1, flow_init, g)
        ops.flow_update(ws.coords1, ws.flow, None, g)
        return att

    def forward(self, image1, image2, iters=12, flow_init=None, upsample=True, test_mode=0):
        """Estimate optical flow between a pair of frames (same contract as core/network.py:164-267)."""
        if not image1.is_cuda:
            raise RuntimeError("craft_b200.CRAFT runs on a CUDA (sm_100a) device only; there is no CPU path")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training (train.py:228, train_ddp.py:246): kernel forward + recompute-in-PyTorch backward for the
            # attention blocks, autograd-native PyTorch for the rest (craft_b200/train_path.py)
            from . import train_path
            with torch.cuda.device(image1.device), ops.precision(self.act_dtype):
                return train_path.forward_train(self, image1.float(), image2.float(), iters, flow_init, test_mode)
        B, _, H, W = image1.shape
        if H % 8 or W % 8:
            raise ValueError("image sides must be multiples of 8 (use InputPadder, as the reference drivers do)")
        # everything below (streams, workspaces, the library's per-device state) refers to the input's device:
        # nn.DataParallel calls each replica from its own thread with another device current
        with torch.cuda.device(image1.device), ops.precision(self.act_dtype):
            savecorr = "SAVECORR" in os.environ      # core/corr.py:180-184 debugging hook: needs the stored volume
            # nn.DataParallel replicas run concurrently in threads: a (global-mode) stream capture in one thread is
            # invalidated by allocations in another, so replicas launch eagerly
            if self.use_cuda_graph and not self.training and not torch.cuda.is_current_stream_capturing() \
                    and not savecorr and not getattr(self, "_is_replica", False):
                return self._forward_graphed(image1, image2, iters, flow_init, test_mode)
            return self._forward_impl(image1, image2, iters, flow_init, test_mode, savecorr=savecorr)

    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in self.buffers())

    def _forward_graphed(self, image1, image2, iters, flow_init, test_mode):
        key = (image1.device.index, tuple(image1.shape), int(iters), int(test_mode), flow_init is not None,
               self.materialize_level0, hp.level0_mode(self.level0), self.encoder_tf32, self.encoder_half, self.elide_dead_upsample, self.precision,
               self._lane)
        sig = self._weights_signature()
        ent = self._graphs.get(key)
        if ent is not None:
            self._graphs.move_to_end(key)
        if ent is None or ent["sig"] != sig:
            ent = dict(sig=sig)
            ent["i1"] = image1.float().clone()
            ent["i2"] = image2.float().clone()
            ent["fi"] = flow_init.float().clone() if flow_init is not None else None
            # warm-up and capture execute the forward three times; the attention diagnostics (max_attn / clamp_count,
            # accumulated on the device) must count this call once
            diags = [t for m in self.modules() if hasattr(m, "_diag") for t in m._diag.values()]
            saved = [t.clone() for t in diags]
            side = torch.cuda.Stream(device=image1.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):      # lazy init (cuDNN autotune, workspaces, packed weights) outside capture
                    self._forward_impl(ent["i1"], ent["i2"], iters, ent["fi"], test_mode)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(image1.device)
            g = torch.cuda.CUDAGraph()
            from . import _lib
            n0 = _lib.launch_count()
            # an explicit capture stream: torch.cuda.graph's default one is a process-wide singleton created on
            # whatever device was current first -- capturing a cuda:1 forward on it invalidates the capture
            with torch.cuda.graph(g, stream=side):
                ent["out"] = self._forward_impl(ent["i1"], ent["i2"], iters, ent["fi"], test_mode)
            ent["graph"] = g
            for m in self.modules():           # diagnostics tensors created during the warm-up start from zero
                if hasattr(m, "_diag"):
                    for t in m._diag.values():
                        if not any(t is d for d in diags):
                            t.zero_()
            for t, sv in zip(diags, saved):
                t.copy_(sv)
            ent["launches"] = _lib.launch_count() - n0   # craft_b200 kernels per replay
            # the graph bakes in the workspace's addresses: keep the buffers alive as long as the graph is
            g8 = TokenGrid(image1.shape[2] // 8, image1.shape[3] // 8)
            ent["ws"] = self._workspaces.get(g8, image1.device, self.materialize_level0 or self.level0, slot=self._ws_slot())
            self._graphs[key] = ent
            while len(self._graphs) > self.max_cached_graphs:
                self._graphs.popitem(last=False)
        ent["i1"].copy_(image1)
        ent["i2"].copy_(image2)
        if flow_init is not None:
            ent["fi"].copy_(flow_init)
        ent["graph"].replay()
        self.launches_last_forward = ent["launches"]
        out = ent["out"]
        # hand out copies: the graph's output buffers are overwritten by the next replay
        if isinstance(out, tuple):
            return tuple([t.clone() for t in o] if isinstance(o, list) else o.clone() for o in out)
        return [t.clone() for t in out]

    def _forward_impl(self, image1, image2, iters, flow_init, test_mode, savecorr=False):
        B, _, H, W = image1.shape
        fmap1, fmap2, cnet_feat = self._encoders(image1.float(), image2.float())
        g = TokenGrid(H // 8, W // 8)
        ws = self._workspaces.get(g, image1.device, self.materialize_level0 or savecorr or self.level0, slot=self._ws_slot())
        dev = image1.device
        flow_lo = torch.empty((B, 2, g.H, g.W), dtype=torch.float32, device=dev)
        n_up = iters if test_mode != 1 else 1
        flow_ups = [torch.empty((B, 2, H, W), dtype=torch.float32, device=dev) for _ in range(n_up)]
        corr_fn = self.corr_fn if self.args.craft else _PlainLookup(self.args.corr_radius)
        for b in range(B):
            fi = flow_init[b].float().contiguous() if flow_init is not None else None
            att = self._prepare_pair(ws, fmap1[b], fmap2[b], cnet_feat[b], fi)
            if savecorr and self.args.craft:
                self.corr_fn._save_volume(ws, os.environ["SAVECORR"])
            main, side = torch.cuda.current_stream(), ws.side
            for itr in range(iters):
                need_up = not (test_mode == 1 and self.elide_dead_upsample) or itr == iters - 1
                lookup = lambda part: corr_fn.lookup_rows(ws, ws.coords1, out_b=ws.CORR, part=part)
                # coords1 += delta / flow = coords1 - coords0 (core/network.py:247,236) happen in the flow head's epilogue
                self.update_block.step(ws, att, itr, need_mask=need_up, lookup=lookup, update_flow=True)
                if not need_up:
                    continue
                # the reference upsamples after every iteration (core/network.py:250-260); nothing in the next
                # iteration depends on it, so it runs on the side stream (mask buffers alternate, hotpath.heads)
                dst = flow_ups[itr if test_mode != 1 else 0][b]
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.upsample_flow(ws.MASKS[itr & 1], ws.flow, g, out=dst)
            main.wait_stream(side)
            ops.unpack_tokens(ws.flow, 0, 2, g, out=flow_lo[b])
        self.call_counter += 1
        if test_mode == 1:
            return flow_lo, flow_ups[0]
        if test_mode == 2:
            return flow_lo, flow_ups
        return flow_ups


class _PlainLookup:
    """Lookup on a pyramid built by hotpath.build_correlation(M=1) (non-craft CorrBlock path)."""

    def __init__(self, radius):
        self.radius, self.num_levels = radius, 4

    lookup_rows = CorrBlock.lookup_rows


RAFTER = CRAFT   # back-compat alias used by evaluate.py:17-19 when unpickling old checkpoints
