"""Development check of the tensor-core Att-Pair scorer (csrc/muav_scorer_tc.cu) on a GPU box: scores against the PyTorch
module and the FP32-pipe kernel, per-stage activations of CTA 0's first pass against the module's intermediates, and the
full-batch launch time of both kernels.  python tools/tc_scorer_check.py [n_envs]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from multi_uav_ta_gym_env_b200 import AllocSpec, BatchedMultiUAVEnv, wps_config  # noqa: E402
from multi_uav_ta_gym_env_b200.scorers import AttPairNet, FusedAttPairScorer, pair_scores  # noqa: E402


def stage_reference(net, tok, envs):
    """rows of CTA 0's first pass -> (stage tensors [4][R][64], row validity)"""
    tf, af = tok["task_feats"], tok["agent_feats"]
    tm, am = tok["task_mask_u8"].bool(), tok["agent_mask_u8"].bool()
    with torch.no_grad():
        t_emb = net.task_proj(tf) + net.type_embed.weight[1]
        a_emb = net.agent_proj(af) + net.type_embed.weight[0]
        x = torch.cat([a_emb, t_emb], 1)
        pad = torch.cat([am, tm], 1)
        layer = net.self_encoder.layers[0]
        sa = layer.self_attn(x, x, x, key_padding_mask=pad, need_weights=False)[0]
        x1 = layer.norm1(x + sa)
        h = layer.norm2(x1 + layer.linear2(torch.relu(layer.linear1(x1))))
        MA = af.shape[1]
        a_h, t_h = h[:, :MA], h[:, MA:]
        a_ctx = net.cross_a2t(a_h, t_h, t_h, key_padding_mask=tm, need_weights=False)[0]
        t_ctx = net.cross_t2a(t_h, a_h, a_h, key_padding_mask=am, need_weights=False)[0]
        z = torch.cat([a_h + a_ctx, t_h + t_ctx], 1)
    stages = [x, x1, h, z]
    # packing of the pass (same rule as the kernel)
    segs, sa_, st_ = [], 0, 0
    for e in envs:
        na = int((~am[e]).sum().item())
        nt = int((~tm[e]).sum().item())
        if na == 0 or nt == 0:
            continue
        if ((sa_ + na + 3) & ~3) + st_ + nt > 128 or len(segs) == 8:
            break
        segs.append((e, sa_, st_, na, nt))
        sa_ += na
        st_ += nt
    split = (sa_ + 3) & ~3
    rows = {}
    for e, ab, tb, na, nt in segs:
        for i in range(na):
            rows[ab + i] = (e, i)
        for j in range(nt):
            rows[split + tb + j] = (e, MA + j)
    return stages, rows


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda")
    cfg = wps_config("WPS_hard")
    env = BatchedMultiUAVEnv(cfg, n, device="cuda:0").reset([s % 512 for s in range(n)])
    env.step_allocated(AllocSpec.local_hungarian(20), n_steps=40)
    tok = env.enable_fused_tokens(32, 16, 15, 0b111)
    env.refresh_fused_tokens()
    torch.manual_seed(1)
    net = AttPairNet().cuda().eval()
    eager_tok = {"task_feats": tok["task_feats"], "task_mask": tok["task_mask_u8"].bool(), "agent_feats": tok["agent_feats"],
                 "agent_mask": tok["agent_mask_u8"].bool(), "edge_valid": tok["edge_valid"]}
    want = pair_scores(net, eager_tok)
    os.environ["MUAV_SCORER_TC"] = "0"
    fp32 = FusedAttPairScorer(net, dev)
    os.environ["MUAV_SCORER_TC"] = "1"
    tc = FusedAttPairScorer(net, dev)
    assert tc.tcw is not None and fp32.tcw is None
    got32 = torch.full_like(want, 7.0)
    fp32.score(tok, got32)
    torch.cuda.synchronize()
    print("fp32 kernel vs torch: max err %.3e" % (got32 - want).abs().max().item(), flush=True)

    # the activation dumps / stage timestamps exist only in a -DMUAV_TC_DEBUG build of the scorer: compile one next to the
    # product library and point the scorer object at it
    import subprocess
    dbg_so = os.path.join(ROOT, "build", "libmuav_tc_debug.so")
    os.makedirs(os.path.dirname(dbg_so), exist_ok=True)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-DMUAV_TC_DEBUG", "-shared", "-o", dbg_so,
                    os.path.join(ROOT, "multi_uav_ta_gym_env_b200", "csrc", "muav_scorer_tc.cu")], check=True)
    dbg_dll = C.CDLL(dbg_so)
    dbg_dll.muav_att_pair_scores_tc.restype = C.c_int
    dbg_dll.muav_att_pair_scores_tc.argtypes = tc.lib.dll.muav_att_pair_scores_tc.argtypes

    class _Lib:
        dll = dbg_dll
    prod_lib, tc.lib = tc.lib, _Lib
    dbg_all = torch.zeros(4 * 128 * 64 + 1024, device=dev)
    dbg = dbg_all[:4 * 128 * 64].view(4, 128, 64)
    dbg.fill_(float("nan"))
    dbg_dll.muav_tc_debug_buffer_(C.c_void_p(dbg_all.data_ptr()))
    got = torch.full_like(want, 7.0)
    tc.score(tok, got)
    torch.cuda.synchronize()
    dbg_dll.muav_tc_debug_buffer_(None)
    tc.lib = prod_lib   # everything below runs the product library
    got_prod = torch.full_like(want, 7.0)
    tc.score(tok, got_prod)
    torch.cuda.synchronize()
    assert torch.equal(got_prod, got), "debug and product builds disagree"
    ts = dbg_all[4 * 128 * 64:].view(torch.int64).cpu().tolist()
    n_ts = ts[0]
    marks = ts[1:1 + n_ts]
    print("stage timestamps of CTA 0 (worker thread 0), cycles since the first: wait_d marks follow arrive marks")
    print(" ".join(str(m - marks[0]) for m in marks), flush=True)
    err = (got - want).abs()
    print("tc kernel vs torch:   max err %.3e   (vs fp32 kernel %.3e)  untouched %d" %
          (err.max().item(), (got - got32).abs().max().item(), int((got == 7.0).sum().item())), flush=True)
    ok = err.max().item() < 2e-5
    group = int(os.environ.get("MUAV_SCORER_TC_GROUP", "0")) or 16
    stages, rows = stage_reference(net, tok, list(range(min(group, n))))
    names = ["x0 = proj + type_embed", "x1 = LN1(x + SA)", "h = LN2(x1 + FF)", "z = h + cross"]
    for s in range(4):
        worst = 0.0
        for r, (e, t) in rows.items():
            worst = max(worst, (dbg[s, r] - stages[s][e, t]).abs().max().item())
        print("stage %d (%s): max err over %d rows %.3e" % (s, names[s], len(rows), worst), flush=True)
    if not ok:
        bad = (err > 2e-5).nonzero()
        print("bad entries:", bad.shape[0], "first:", bad[:8].tolist())
        per_env = (err.flatten(1).max(1).values > 2e-5).nonzero().flatten()
        print("bad envs:", per_env.numel(), per_env[:16].tolist())
        sys.exit(1)

    # subset / need paths
    idx = torch.arange(3, min(n, 150), 2, device=dev, dtype=torch.int32)
    g2 = torch.full_like(want, 7.0)
    tc.score(tok, g2, idx)
    assert (g2[idx.long()] - want[idx.long()]).abs().max().item() < 2e-5 and bool((g2[0] == 7.0).all())

    def timeit(f, reps=20):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    print("full batch (%d envs): fp32 kernel %.4f ms, tc kernel %.4f ms" %
          (n, timeit(lambda: fp32.score(tok, got32)), timeit(lambda: tc.score(tok, got))), flush=True)
    need = tok["need"]
    need.zero_()
    need[torch.arange(0, n, 4, device=dev)] = 1
    print("need = every 4th env:  fp32 kernel %.4f ms, tc kernel %.4f ms" %
          (timeit(lambda: fp32.score(tok, got32, use_need=True)), timeit(lambda: tc.score(tok, got, use_need=True))), flush=True)
    for g in (6, 8, 0):
        os.environ["MUAV_SCORER_TC_GROUP"] = str(g)
        print("  group %d: full %.4f ms, need/4 %.4f ms" %
              (g, timeit(lambda: tc.score(tok, got)), timeit(lambda: tc.score(tok, got, use_need=True))), flush=True)
    print("TC CHECK OK")


if __name__ == "__main__":
    main()
