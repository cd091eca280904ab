"""Bring-up diagnostics (run on the GPU box): each stage in its own process so a trap cannot hide the rest.
usage: python tools/gpu_diag.py <stage>     stages: info gemm attn attn_variants
The timeline stages (attn_trace, attn_cta) read B2_ATT_TRACE / B2_ATT_CTATRACE, which only a diagnostics build honours:
    B2_DEFINES=-DB2_DIAG python llm_np_cp_b200/build.py --force
and they trace round 1's one-tile-per-CTA attention kernel (run them with B2_ATTN_CO_ONESHOT=1)."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def stage_info():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), torch.version.cuda)
    p = torch.cuda.get_device_properties(0)
    print("SMs", p.multi_processor_count, "mem GB", p.total_memory / 2**30)
    from llm_np_cp_b200 import _lib
    print("b2_device_count", _lib.lib().b2_device_count())


def stage_gemm():
    from gpu_util import bf16_randn, op_gemm
    for (M, N, K, bn) in [(128, 256, 64, 256), (128, 256, 128, 256), (128, 64, 64, 64), (256, 512, 256, 256), (4096, 4096, 4096, 256)]:
        a, w = bf16_randn(M, K, seed=1), bf16_randn(N, K, seed=2)
        out = op_gemm(a, w, epi=1, bn=bn)
        ref = a.float() @ w.float().T
        d = (out - ref).abs()
        print(f"gemm M{M} N{N} K{K} bn{bn}: max_abs {d.max().item():.4g} ref_absmax {ref.abs().max().item():.4g} "
              f"out_absmax {out.abs().max().item():.4g} frac_bad {(d > 1e-2).float().mean().item():.4f}")
        if d.max() > 1e-2:
            bad = (d > 1e-2)
            print("  bad rows (first 16):", torch.nonzero(bad.any(1)).flatten()[:16].tolist())
            print("  bad cols (first 16):", torch.nonzero(bad.any(0)).flatten()[:16].tolist())
            print("  out[0,:8]", out[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
            # is it a K-slice problem?  compare against partial-K references
            for kk in range(16, K + 1, 16):
                r2 = a[:, :kk].float() @ w[:, :kk].float().T
                if (out - r2).abs().max() < 1e-2:
                    print(f"  matches partial K={kk}")
    # timing of a big one
    M = N = K = 8192
    a, w = bf16_randn(M, K, seed=1), bf16_randn(N, K, seed=2)
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        op_gemm(a, w, epi=0, out=out)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        op_gemm(a, w, epi=0, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"gemm 8192^3 bf16: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    ref = a[:256].float() @ w.float().T
    print("  8192 check max_abs", (out[:256].float() - ref).abs().max().item(), "ref absmax", ref.abs().max().item())


def attn_case(B, S, NH, NKV, HD, t0=0, causal=1, window=0, softcap=0.0):
    from gpu_util import bf16_randn, op_attention, ref_attention
    Tmax = t0 + S + 3
    q = bf16_randn(B * S, NH * HD, seed=13)
    kc, vc = bf16_randn(B, NKV, Tmax, HD, seed=14), bf16_randn(B, NKV, Tmax, HD, seed=15)
    scale = 1.0 / math.sqrt(HD)
    out = op_attention(q, kc, vc, B, S, NH, NKV, HD, Tmax, t0, scale, softcap, causal, window)
    ref = ref_attention(q, kc, vc, B, S, NH, NKV, HD, t0, scale, softcap, causal, window)
    d = (out.float() - ref).abs()
    return d.max().item(), d.mean().item(), out, ref


def stage_attn():
    for case in [(1, 128, 1, 1, 64), (1, 16, 1, 1, 64), (1, 300, 2, 1, 64), (1, 128, 1, 1, 128), (1, 128, 1, 1, 256), (2, 2048, 8, 2, 64)]:
        mx, mean, out, ref = attn_case(*case)
        print(f"attn {case}: max_abs {mx:.4g} mean_abs {mean:.4g} out_absmax {out.float().abs().max().item():.4g} ref_absmax {ref.abs().max().item():.4g}")
        if mx > 5e-2:
            print("   out[0,:8]", out[0, :8].float().tolist())
            print("   ref[0,:8]", ref[0, :8].tolist())
            print("   out[5,:8]", out[5, :8].float().tolist())
            print("   ref[5,:8]", ref[5, :8].tolist())


def stage_attn_variants():
    for HD, bnkv in [(64, 128), (128, 128)]:
        for lbo, sbo in [(0, 0), (1024, bnkv * 128), (128, 1024), (1024, 128), (bnkv * 128, 128), (16, 1024), (2048, 1024)]:
            os.environ["B2_V_LBO"], os.environ["B2_V_SBO"] = str(lbo), str(sbo)
            try:
                mx, mean, _, _ = attn_case(1, 128, 1, 1, HD)
                print(f"HD{HD} v_lbo={lbo} v_sbo={sbo}: max_abs {mx:.4g} mean_abs {mean:.4g}")
            except Exception as e:
                print(f"HD{HD} v_lbo={lbo} v_sbo={sbo}: EXC {e}")
                return




def stage_model():
    from oracle import np_oracle as O
    import llm_np_cp_b200 as M
    cfg = O.llama_config("tiny")
    w = O.synth_weights(cfg, 1, 0.08, 0.1)
    Wn = O.to_numpy_f32(w)
    model = M.LlamaForCausalLM(M.AttributeDict(cfg), w)
    for (B, S) in [(2, 12), (1, 12), (1, 2), (1, 9), (2, 5), (1, 16), (1, 200), (2, 12)]:
        ids = np.random.default_rng(B * 100 + S).integers(0, cfg.vocab_size, (B, S))
        want = O.forward(cfg, Wn, ids)
        _, lg, _, hs, _ = model(ids, use_cache=False)
        errs = [float(np.abs(h.cpu().numpy() - wh).max()) for h, wh in zip(hs, want.hidden_states)]
        d = np.abs(lg.cpu().numpy() - want.logits)
        rows = d.reshape(B * S, -1).max(1)
        print(f"B{B} S{S}: logits max {d.max():.4g}; hidden max errs {['%.3g' % e for e in errs]}; bad rows {np.nonzero(rows > 0.3)[0][:20].tolist()}")
        h1 = hs[1].cpu().numpy().reshape(B * S, -1)
        w1 = want.hidden_states[1].reshape(B * S, -1)
        print("    layer-1 row errs:", ["%.2g" % e for e in np.abs(h1 - w1).max(1)[:24]])


def stage_fresh():
    from oracle import np_oracle as O
    import llm_np_cp_b200 as M
    cfg = O.llama_config("tiny")
    w = O.synth_weights(cfg, 1, 0.08, 0.1)
    Wn = O.to_numpy_f32(w)
    keep = []
    for (B, S) in [(1, 2), (1, 9), (2, 12), (1, 200), (2, 5), (1, 12), (2, 2)]:
        model = M.LlamaForCausalLM(M.AttributeDict(cfg), w)
        keep.append(model)
        ids = np.random.default_rng(0).integers(0, cfg.vocab_size, (B, S))
        want = O.forward(cfg, Wn, ids)
        for rep in range(2):
            _, lg, _, hs, _ = model(ids, use_cache=False)
            errs = [float(np.abs(h.cpu().numpy() - wh).max()) for h, wh in zip(hs, want.hidden_states)]
            d = np.abs(lg.cpu().numpy() - want.logits)
            h1 = hs[1].cpu().numpy().reshape(B * S, -1)
            w1 = want.hidden_states[1].reshape(B * S, -1)
            print(f"fresh B{B} S{S} rep{rep}: logits max {d.max():.4g}; hidden errs {['%.3g' % e for e in errs]}; "
                  f"layer1 row errs {['%.2g' % e for e in np.abs(h1 - w1).max(1)[:12]]}")


def stage_destroy():
    import gc
    from oracle import np_oracle as O
    import llm_np_cp_b200 as M
    from conftest import load_golden
    cfg = O.llama_config("tiny")
    w = O.synth_weights(cfg, 1, 0.08, 0.1)
    Wn = O.to_numpy_f32(w)

    def run(B, S, tag):
        model = M.LlamaForCausalLM(M.AttributeDict(cfg), w)
        ids = np.random.default_rng(0).integers(0, cfg.vocab_size, (B, S))
        want = O.forward(cfg, Wn, ids)
        _, lg, _, hs, _ = model(ids, use_cache=False)
        errs = [float(np.abs(h.cpu().numpy() - wh).max()) for h, wh in zip(hs, want.hidden_states)]
        d = np.abs(lg.cpu().numpy() - want.logits)
        print(f"{tag} B{B} S{S}: logits max {d.max():.4g}; hidden errs {['%.3g' % e for e in errs]}", flush=True)
        return model

    m = run(2, 12, "first")
    del m
    gc.collect()
    print("destroyed engine 1", flush=True)
    m2 = run(1, 2, "after-destroy")
    m3 = run(1, 9, "after-destroy(no destroy between)")
    del m2, m3
    gc.collect()
    m4 = run(2, 12, "after-destroy-2")
    m5 = run(1, 2, "no-destroy")
    # golden flow
    for name in ["llama_tiny_prefill", "llama_tiny_qlen2"]:
        meta, arr = load_golden(name)
        model = M.LlamaForCausalLM(M.AttributeDict(cfg), O.synth_weights(cfg, meta["wseed"], meta["std"], meta["jitter"]))
        _, lg, _, hs, _ = model(arr["ids"], use_cache=False)
        d = np.abs(lg.cpu().numpy() - arr["logits"])
        print(f"golden {name}: logits max {d.max():.4g}", flush=True)
        del model
        gc.collect()


def stage_attn_perf():
    import ctypes
    from gpu_util import bf16_randn, L, stream
    from llm_np_cp_b200 import _lib
    for (B, S, NH, NKV, HD) in [(8, 2048, 32, 8, 64), (2, 2048, 32, 8, 64), (1, 8192, 24, 8, 128), (2, 4096, 8, 4, 256)]:
        Tmax = S
        q = bf16_randn(B * S, NH * HD, seed=13)
        kc, vc = bf16_randn(B, NKV, Tmax, HD, seed=14), bf16_randn(B, NKV, Tmax, HD, seed=15)
        out = torch.zeros((B * S, NH * HD), dtype=torch.bfloat16, device="cuda")
        scale = 1.0 / math.sqrt(HD)
        def run():
            _lib.check(L().b2_op_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), B, S, NH, NKV, HD,
                                           Tmax, 0, float(scale), 0.0, 1, 0, 0, stream()))
        run(); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        n = 5
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4.0 * B * NH * HD * S * (S + 1) / 2
        print(f"attn B{B} S{S} NH{NH} NKV{NKV} HD{HD}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (causal flops)", flush=True)


def stage_attn_trace():
    """clock64 timeline of CTA (0,0,0) of the pair kernel (heaviest q-tile pair) while the whole grid runs."""
    from gpu_util import bf16_randn, L, stream
    from llm_np_cp_b200 import _lib
    B, S, NH, NKV, HD = 8, 2048, 32, 8, 64
    q = bf16_randn(B * S, NH * HD, seed=13)
    kc, vc = bf16_randn(B, NKV, S, HD, seed=14), bf16_randn(B, NKV, S, HD, seed=15)
    out = torch.zeros((B * S, NH * HD), dtype=torch.bfloat16, device="cuda")
    trace = torch.zeros(3 * 1024, dtype=torch.int64, device="cuda")
    os.environ["B2_ATT_TRACE"] = hex(trace.data_ptr())
    for _ in range(2):
        trace.zero_()
        _lib.check(L().b2_op_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), B, S, NH, NKV, HD, S, 0,
                                       float(1 / math.sqrt(HD)), 0.0, 1, 0, 0, stream()))
        torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(3, 512, 2)
    t00 = min(int(t[r, 0, 1]) for r in range(3) if t[r, 0, 1] > 0)
    names = {0: "mma:start", 1: "mma:q_ready", 10: "mma:sfree0", 11: "mma:sfree1", 12: "mma:k_next0", 13: "mma:k_next1", 20: "mma:pready0",
             21: "mma:pready1", 22: "mma:pv0_issued", 23: "mma:pv1_issued", 30: "sm:wait_s", 31: "sm:s_full", 32: "sm:s_in_regs",
             33: "sm:max_done", 34: "sm:o_done", 35: "sm:p_written"}
    for r, nm in enumerate(["MMA", "SOFTMAX g0", "SOFTMAX g1"]):
        ev = [(int(a), int(c) - t00) for a, c in t[r] if c > 0]
        print(f"--- {nm}: {len(ev)} events, span {ev[-1][1] if ev else 0} cycles")
        prev = 0
        for k, (a, c) in enumerate(ev[:70] if r == 0 else ev[:44]):
            print(f"   {c:8d} (+{c - prev:6d}) {names.get(a, a)}")
            prev = c
        if r > 0 and ev:
            # per-tile summary over all tiles
            import collections
            d = collections.defaultdict(list)
            last = {}
            for a, c in ev:
                last[a] = c
                if a == 31 and 30 in last: d["wait s_full"].append(c - last[30])
                if a == 32 and 31 in last: d["ld S -> regs"].append(c - last[31])
                if a == 33 and 32 in last: d["row max"].append(c - last[32])
                if a == 34 and 33 in last: d["wait o_done"].append(c - last[33])
                if a == 37: d["exp loop"].append(c - last.get(34, last.get(33, c)))
                if a == 38 and 37 in last: d["fence.proxy"].append(c - last[37])
                if a == 35 and 38 in last: d["arrive p_ready"].append(c - last[38])
            for k2, v in d.items():
                print(f"   [{nm}] {k2:14s} mean {np.mean(v):8.1f}  max {max(v):6d}  n={len(v)}")


def stage_attn_cta():
    """Per-CTA {start, end, smid} of the pair kernel: where the kernel's wall time goes (busy vs gaps vs tail)."""
    from gpu_util import bf16_randn, L, stream
    from llm_np_cp_b200 import _lib
    for (B, S, NH, NKV, HD) in [(8, 2048, 32, 8, 64), (1, 8192, 24, 8, 128)]:
        q = bf16_randn(B * S, NH * HD, seed=13)
        kc, vc = bf16_randn(B, NKV, S, HD, seed=14), bf16_randn(B, NKV, S, HD, seed=15)
        out = torch.zeros((B * S, NH * HD), dtype=torch.bfloat16, device="cuda")
        npair = S // 128                   # q tiles = blockIdx.x range (heaviest first)
        ncta = npair * NH * B
        tb = torch.zeros(ncta * 4, dtype=torch.int64, device="cuda")
        os.environ["B2_ATT_CTATRACE"] = hex(tb.data_ptr())
        for _ in range(3):
            tb.zero_()
            _lib.check(L().b2_op_attention(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), out.data_ptr(), B, S, NH, NKV, HD, S, 0,
                                           float(1 / math.sqrt(HD)), 0.0, 1, 0, 0, stream()))
            torch.cuda.synchronize()
        del os.environ["B2_ATT_CTATRACE"]
        t = tb.cpu().numpy().reshape(ncta, 4)
        st, en, sm = t[:, 0], t[:, 1], t[:, 2]
        T0, T1 = st.min(), en.max()
        dur = (en - st).astype(np.float64)
        print(f"== attn B{B} S{S} NH{NH} HD{HD}: {ncta} CTAs, kernel span {(T1 - T0) / 1e3:.1f} us, sum(CTA time)/148 = {dur.sum() / 148 / 1e3:.1f} us")
        x = np.arange(ncta) % npair           # blockIdx.x (pair index reversed: 0 = heaviest)
        for bx in range(npair):
            d = dur[x == bx]
            iters = npair - bx
            print(f"   blockIdx.x {bx:2d} ({iters:3d} kv iters): mean {d.mean() / 1e3:7.2f} us  min {d.min() / 1e3:7.2f}  max {d.max() / 1e3:7.2f}")
        gaps, busy, last_end, ghz = [], [], [], []
        for s_ in np.unique(sm):
            idx = np.where(sm == s_)[0]
            o = idx[np.argsort(st[idx])]
            if len(o) > 1 and en[o[-1]] > en[o[0]]:
                ghz.append((t[o[-1], 3] - t[o[0], 3]) / float(en[o[-1]] - en[o[0]]))
            busy.append(dur[o].sum())
            last_end.append(en[o].max())
            gaps += list((st[o][1:] - en[o][:-1]).astype(np.float64))
            gaps.append(float(st[o][0] - T0))
        gaps = np.array(gaps)
        print(f"   SM clock from clock64/globaltimer: {np.median(ghz):.3f} GHz")
        print(f"   SMs used {len(busy)}; busy/span mean {np.mean(busy) / (T1 - T0):.3f} min {np.min(busy) / (T1 - T0):.3f}; "
              f"inter-CTA gap mean {gaps.mean():.0f} ns, p50 {np.median(gaps):.0f}, max {gaps.max():.0f}, total/SM {gaps.sum() / len(busy) / 1e3:.1f} us; "
              f"tail (span - SM's last end) mean {np.mean(T1 - np.array(last_end)) / 1e3:.1f} us")


def stage_gemm2():
    """1-CTA vs 2-CTA GEMM on the model's prefill shapes (event-timed, inputs >> L2 between variants)."""
    from gpu_util import bf16_randn, op_gemm
    for (M, N, K, name) in [(16384, 16384, 2048, "gate/up"), (16384, 2048, 8192, "down"), (16384, 3072, 2048, "qkv-shape"), (16384, 2048, 2048, "o_proj")]:
        a, w = bf16_randn(M, K, seed=1), bf16_randn(N, K, seed=2, std=0.05)
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        ref = a[:256].float() @ w.float().T
        for bn in (256, 512):
            for _ in range(2):
                op_gemm(a, w, epi=0, bn=bn, out=out)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            n = 10
            for _ in range(n):
                op_gemm(a, w, epi=0, bn=bn, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            err = (out[:256].float() - ref).abs().max().item()
            print(f"{name:10s} M{M} N{N} K{K} {'2-CTA' if bn == 512 else '1-CTA'}: {ms * 1e3:8.1f} us  {2 * M * N * K / ms / 1e9:7.1f} TFLOP/s  max_err {err:.3g}", flush=True)


if __name__ == "__main__":
    globals()["stage_" + sys.argv[1]]()
