"""Launch one hot-path kernel a few times on realistic buffers (448x1024 grid) so that
`ncu --set full -k regex:<name>` can capture it without profiling a whole forward.
usage: python profiles/kernel_only.py {pv|pv_f2|corr|lse|lse_f2|gru_zr|lookup0|lookup|heads}[,more] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import H, ITERS, W, _pairs, _state_dict  # noqa: E402
from craft_b200 import hotpath as hp, ops  # noqa: E402
from craft_b200.ops import TokenGrid  # noqa: E402
from craft_b200.setrans import get_workspace  # noqa: E402


def main(which, reps):
    trace = os.environ.pop("CRAFT_PV_TRACE", None)     # only the kernel under test is traced, not the warm-up forward
    gtrace = os.environ.pop("CRAFT_GEMM_TRACE", None)
    dev = torch.device("cuda", 0)
    model, _ = _state_dict()
    model = model.to(dev).eval()
    a, b = _pairs(1, dev)[0]
    with torch.no_grad(), ops.precision(model.act_dtype):
        model(a, b, iters=1, test_mode=1)       # fills every workspace buffer with realistic data
        torch.cuda.synchronize()
        g = TokenGrid(H // 8, W // 8)
        ws = model.workspace_for(8 * g.H, 8 * g.W, dev)
        ub = model.update_block
        uw = ub.weights(g)
        att_tbl = model.att.vispos_encoder.table()
        f2_tbl = model.f2_trans.vispos_encoder.table()
        if os.environ.get("KO_NOBIAS") == "1":      # experiment: no positional bias -> no "near" slow path
            att_tbl = f2_tbl = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def run(which=which):
            if which == "pv":
                ks = ws.pv_split(4)
                ops.attn_pv(ws.Qa, ws.Ka, ws.Vt, g, M=4, d=32, F=128, w_pos=1.0, pos_table=att_tbl, R=7,
                            clip=ws.clip_att, lse2=ws.lse2_att, out=ws.opart(ks, 4, 128), ksplit=ks, zero_fill=False)
            elif which == "pv_f2":
                ks = ws.pv_split(4)
                ops.attn_pv(ws.Q2, ws.K2, ws.Vt, g, M=4, d=64, F=256, w_pos=0.5, pos_table=f2_tbl, R=7,
                            clip=ws.clip_f2, lse2=ws.lse2_f2, out=ws.opart(ks, 4, 256), ksplit=ks, zero_fill=False)
            elif which == "corr":
                ws.stat_sum.zero_()
                ops.corr_build(ws.Qc, ws.Kc, g, M=4, d=64, w_agg=0.1285, w_pos=0.5, pos_table=f2_tbl, R=7,
                               clip=ws.inf_clip, stat_sum=ws.stat_sum[0], stat_max=ws.stat_max[0:1], levels=ws.levels,
                               ksplit=ws.ks_sc, level0_h16=ws.level0_h16)
            elif which == "lse_f2":
                ops.attn_lse(ws.Q2, ws.K2, g, M=4, d=64, w_pos=0.5, pos_table=f2_tbl, R=7, clip=ws.inf_clip,
                             stat_max=ws.stat_max[1:2], lse_part=ws.lse_part, lse2=ws.lse2_f2, ksplit=ws.ks_sc)
            elif which == "lse":
                ops.attn_lse(ws.Qa, ws.Ka, g, M=4, d=32, w_pos=1.0, pos_table=att_tbl, R=7, clip=ws.inf_clip,
                             stat_max=ws.stat_max[2:3], lse_part=ws.lse_part, lse2=ws.lse2_att, ksplit=ws.ks_sc)
            elif which in ("gru_zr", "gru_q"):
                hp.sep_conv_gru(ws, uw)
            elif which == "lookup0":
                ops.corr_lookup0(grid=g, coords=ws.coords1, mean_rstd=ws.mean_rstd, out_b=ws.CORR, **ws.corr_meta)
            elif which == "lookup":
                ops.corr_lookup(ws.levels, g, ws.coords1, ws.mean_rstd, out_b=ws.CORR, first_level=1)
            elif which == "lookup_all":
                ops.corr_lookup(ws.levels, g, ws.coords1, ws.mean_rstd, out_b=ws.CORR, level0_h16=ws.level0_h16)
            elif which == "finalize":
                agg = ub.aggregator.packed()
                ks = ws.pv_split(4)
                ops.modes_finalize(ws.opart(ks, 4, 128), ks, 4, 128, g, w_score=agg["ws"], b_score=agg["bs"], coeff=agg["coeff"],
                                   x_b=ws.X, colx=256, out_b=ws.X, colb=384, pv_bk=ops.pv_block_keys(32, 128))
            elif which == "heads":
                hp.heads(ws, uw)
            elif which == "iter_gemms":     # every tensor-core GEMM of one refinement iteration but the V^T one
                hp.motion_encoder(ws, uw, None)
                hp.sep_conv_gru(ws, uw)
                hp.heads(ws, uw, 0, need_mask=False)
            else:
                raise SystemExit("unknown kernel " + which)

        if trace:
            os.environ["CRAFT_PV_TRACE"] = trace
        if gtrace:
            os.environ["CRAFT_GEMM_TRACE"] = gtrace
        prof = os.environ.get("KO_PROFILE") == "1"      # ncu --profile-from-start off: only the measured calls are profiled
        for w in which.split(","):          # several kernels in one process: "pv,corr,lse"
            run(w)
            torch.cuda.synchronize()
            if prof:
                torch.cuda.profiler.start()
            e0.record()
            for _ in range(reps):
                run(w)
            e1.record()
            torch.cuda.synchronize()
            if prof:
                torch.cuda.profiler.stop()
            print("%s: %.2f us per call (avg of %d)" % (w, 1000 * e0.elapsed_time(e1) / reps, reps), flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 5)
