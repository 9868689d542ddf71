"""Operator-level GPU probe: runs each kernel family against a plain torch reference on the B200 and prints
error statistics. Development aid (the judged parity tests are tests/test_gpu_*.py against oracle/).

usage:  python tools/gpu_probe.py all            # every case, each in its own subprocess with a timeout
        python tools/gpu_probe.py <case-name>    # one case in this process
"""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ishara_b200 import _lib

torch.manual_seed(0)
dev = "cuda:0"


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    mx = err.max().item()
    bad = (err > tol * (1 + ref.abs())).float().mean().item()
    ok = bad == 0.0 and bool(torch.isfinite(got).all())
    print(f"[{ 'PASS' if ok else 'FAIL'}] {name}: max_abs_err={mx:.4g} ref_max={denom:.4g} frac_bad={bad:.4g} tol={tol}")
    if not ok:
        idx = torch.nonzero(err > tol * (1 + ref.abs()))[:8]
        for i in idx:
            i = tuple(i.tolist())
            print("   first bad", i, "got", got[i].item(), "ref", ref[i].item())
        # structure of the error: per 8-row / 8-col block pattern
        if got.dim() == 2:
            e = (err > tol * (1 + ref.abs())).float()
            print("   bad frac by row%8:", [round(e[r::8].mean().item(), 3) for r in range(8)])
            print("   bad frac by col block of 8 (first 16):", [round(e[:, c * 8:(c + 1) * 8].mean().item(), 3) for c in range(min(16, got.shape[1] // 8))])
            print("   bad frac by row tile of 128 (first 8):", [round(e[r * 128:(r + 1) * 128].mean().item(), 3) for r in range(min(8, (got.shape[0] + 127) // 128))])
    return ok


def gemm_case(M, K, N, block_n, act=0, bias=False, gate=False, rowtab=False, resid=False, ln0=False, ln1=False,
              row_mode=False, out_f32=False, nout=0, T=96, time_it=False):
    lib = _lib.load()
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    wt = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    nfull = N // 2 if act == 3 else N
    no = nout if nout else nfull
    bias_t = torch.randn(N, device=dev) if bias else None
    nseq = (M + T - 1) // T
    gate_t = torch.rand(nseq, N, device=dev) + 0.5 if gate else None
    rowtab_t = torch.randn(T, N, device=dev) if rowtab else None
    resid_t = torch.randn(M, no, device=dev).bfloat16() if resid else None
    g0 = torch.rand(N, device=dev) + 0.5 if ln0 else None
    b0 = torch.randn(N, device=dev) if ln0 else None
    g1 = torch.rand(N, device=dev) + 0.5 if ln1 else None
    b1 = torch.randn(N, device=dev) if ln1 else None
    out0 = torch.full((M, no), float("nan"), device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    out1 = torch.full((M, no), float("nan"), device=dev, dtype=torch.bfloat16) if ln1 else None
    args = _lib.GemmArgs()
    args.a = a.data_ptr(); args.wt = wt.data_ptr(); args.out0 = out0.data_ptr()
    args.out1 = out1.data_ptr() if out1 is not None else None
    args.bias = bias_t.data_ptr() if bias else None
    args.gate = gate_t.data_ptr() if gate else None
    args.rowtab = rowtab_t.data_ptr() if rowtab else None
    args.resid = resid_t.data_ptr() if resid else None
    args.ln0_g = g0.data_ptr() if ln0 else None
    args.ln0_b = b0.data_ptr() if ln0 else None
    args.ln0_eps = 1e-3
    args.ln1_g = g1.data_ptr() if ln1 else None
    args.ln1_b = b1.data_ptr() if ln1 else None
    args.ln1_eps = 1e-6
    args.M, args.N, args.K, args.lda, args.nout = M, N, K, K, nout
    args.rows_per_seq = T
    args.act = act
    args.block_n = block_n
    args.out_f32 = int(out_f32)
    args.row_mode = int(row_mode)
    _lib.check(lib.ishara_op_gemm(C.byref(args), None))
    torch.cuda.synchronize()
    # reference
    v = a.float() @ wt.float().t()
    rows = torch.arange(M, device=dev)
    if bias: v = v + bias_t
    if gate: v = v * gate_t[rows // T]
    if rowtab: v = v + rowtab_t[rows % T]
    if act == 1: v = v * torch.sigmoid(v)
    elif act == 2: v = torch.relu(v)
    elif act == 3:
        bn = block_n
        vv = v.view(M, N // bn, 2, bn // 2)
        v = (vv[:, :, 0] * torch.sigmoid(vv[:, :, 1])).reshape(M, N // 2)
    v = v[:, :no]
    if resid: v = v + resid_t.float()
    if ln0: v = torch.nn.functional.layer_norm(v, (N,), g0, b0, 1e-3)
    name = f"gemm M{M} K{K} N{N} bn{block_n} act{act} b{int(bias)} g{int(gate)} rt{int(rowtab)} r{int(resid)} ln{int(ln0)}{int(ln1)} row{int(row_mode)} f32{int(out_f32)}"
    ok = report(name + " out0", out0, v, 2e-2 if not out_f32 else 5e-3)
    if ln1:
        xn = torch.nn.functional.layer_norm(v, (N,), g1, b1, 1e-6)
        ok &= report(name + " out1", out1, xn, 3e-2)
    if time_it:
        for _ in range(3): lib.ishara_op_gemm(C.byref(args), None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the library launches on the legacy default stream when stream == NULL
        s = torch.cuda.default_stream()
        e0.record(s)
        for _ in range(10): lib.ishara_op_gemm(C.byref(args), None)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   time {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
    return ok


def dw_case(B, T, Cc, k, pad_left, post, bias=True, colsum=False):
    lib = _lib.load()
    x = torch.randn(B, T, Cc, device=dev).bfloat16()
    w = torch.randn(k, Cc, device=dev) / k ** 0.5
    b = torch.randn(Cc, device=dev) if bias else None
    eca = torch.randn(5, device=dev) if post == 2 else None
    out = torch.full((B, T, Cc), float("nan"), device=dev, dtype=torch.bfloat16)
    cs = torch.zeros(B, Cc, device=dev) if colsum else None
    _lib.check(lib.ishara_op_dwconv(ptr(x), ptr(out), ptr(w), ptr(b), ptr(eca), ptr(cs), B, T, Cc, k, pad_left, post, None))
    torch.cuda.synchronize()
    xp = torch.nn.functional.pad(x.float().transpose(1, 2), (pad_left, k - 1 - pad_left))
    y = torch.nn.functional.conv1d(xp, w.t().unsqueeze(1).contiguous(), b, groups=Cc).transpose(1, 2)
    if post == 1: y = y * torch.sigmoid(y)
    if post == 2:
        mean = y.mean(1)
        z = torch.nn.functional.conv1d(mean.unsqueeze(1), eca.view(1, 1, 5), padding=2).squeeze(1)
        y = y * torch.sigmoid(z).unsqueeze(1)
    ok = report(f"dwconv B{B} T{T} C{Cc} k{k} pad{pad_left} post{post}", out.view(B * T, Cc), y.reshape(B * T, Cc), 2e-2)
    if colsum:
        ok &= report("   colsum", cs, out.float().sum(1), 1e-3)
    return ok


def attn_case(B, T, H, dh, mask=False, time_it=False):
    lib = _lib.load()
    D = H * dh
    qkv = torch.randn(B * T, 3 * D, device=dev).bfloat16()
    out = torch.full((B * T, D), float("nan"), device=dev, dtype=torch.bfloat16)
    km = None
    if mask:
        km = (torch.rand(B, T, device=dev) > 0.3).to(torch.uint8)
        km[:, 0] = 1
    scale = D ** -0.5
    _lib.check(lib.ishara_op_attention(ptr(qkv), ptr(out), ptr(km), B, T, H, dh, scale, None))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, T, H, 3 * dh).permute(0, 2, 1, 3).split(dh, dim=-1)
    att = q @ k.transpose(-1, -2) * scale
    if mask: att = att + (1 - km.float())[:, None, None, :] * -1e9
    o = (torch.softmax(att, -1) @ v).permute(0, 2, 1, 3).reshape(B * T, D)
    ok = report(f"attention B{B} T{T} H{H} dh{dh} mask{int(mask)}", out, o, 2e-2)
    if time_it:
        for _ in range(3): lib.ishara_op_attention(ptr(qkv), ptr(out), ptr(km), B, T, H, dh, scale, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.default_stream()
        e0.record(s)
        for _ in range(10): lib.ishara_op_attention(ptr(qkv), ptr(out), ptr(km), B, T, H, dh, scale, None)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"   time {ms*1e3:.1f} us  {4.0*B*T*T*D/ms/1e9:.1f} TFLOP/s")
    return ok


def ctc_case(B, T, V, L):
    lib = _lib.load()
    logits = torch.randn(B, T, V, device=dev) * 2
    blank = V - 1
    labels = torch.full((B, L), blank, dtype=torch.int32, device=dev)
    lens = torch.randint(0, L + 1, (B,))
    lens[0] = L
    if B > 1: lens[1] = 0
    for b in range(B):
        labels[b, :lens[b]] = torch.randint(0, V - 1, (int(lens[b]),), dtype=torch.int32)
        if lens[b] > 3: labels[b, 1] = labels[b, 0]  # repeated label
    nll = torch.zeros(B, device=dev)
    grad = torch.zeros(B, T, V, device=dev)
    _lib.check(lib.ishara_ctc_loss(ptr(logits), ptr(labels), B, T, V, L, blank, ptr(nll), ptr(grad), None))
    torch.cuda.synchronize()
    lg = logits.double().detach().cpu().requires_grad_(True)
    lp = torch.log_softmax(lg, -1).transpose(0, 1)
    ref = torch.nn.functional.ctc_loss(lp, labels.cpu().long(), torch.full((B,), T, dtype=torch.long), lens.long(),
                                       blank=blank, reduction="none", zero_infinity=False)
    ref.sum().backward()
    ok = report(f"ctc nll B{B} T{T} V{V} L{L}", nll.cpu(), ref.detach().float(), 1e-3)
    ok &= report("   ctc grad", grad.cpu().view(B * T, V), lg.grad.float().view(B * T, V), 2e-3)
    return ok


def decode_case(B, T, V):
    lib = _lib.load()
    logits = torch.randn(B, T, V, device=dev)
    # make runs: repeat frames
    idx = torch.randint(0, T, (B, T), device=dev).sort(1).values
    logits = torch.gather(logits, 1, idx.unsqueeze(-1).expand(-1, -1, V)).contiguous()
    logits[:, :, V - 1] += 1.0
    ids = torch.zeros(B, T, dtype=torch.int32, device=dev)
    lens = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.check(lib.ishara_greedy_decode(ptr(logits), B, T, V, V - 1, ptr(ids), ptr(lens), None))
    torch.cuda.synchronize()
    ok = True
    am = logits.argmax(-1).cpu()
    for b in range(B):
        x = am[b]
        keep = x[:-1] != x[1:]
        y = x[:-1][keep]
        y = y[y != V - 1]
        got = ids[b, :lens[b]].cpu().long()
        if got.numel() != y.numel() or not torch.equal(got, y):
            ok = False
            print("   decode mismatch seq", b, got.tolist()[:10], y.tolist()[:10])
    print(f"[{'PASS' if ok else 'FAIL'}] decode B{B} T{T} V{V} (mean len {lens.float().mean().item():.1f})")
    return ok


CASES = {
    "gemm_min": lambda: gemm_case(256, 64, 256, 256),
    "gemm_k256": lambda: gemm_case(384, 256, 256, 256),
    "gemm_wide_swish": lambda: gemm_case(1000, 256, 512, 256, act=1, bias=True),
    "gemm_glu": lambda: gemm_case(640, 256, 512, 256, act=3, bias=True),
    "gemm_bn128": lambda: gemm_case(640, 192, 384, 128, act=2, bias=True, resid=True),
    "gemm_row_resid_ln1": lambda: gemm_case(768, 512, 256, 256, bias=True, resid=True, ln1=True, row_mode=True),
    "gemm_row_all": lambda: gemm_case(900, 256, 256, 256, bias=True, gate=True, rowtab=True, resid=True, ln0=True, ln1=True, row_mode=True),
    "gemm_row_plain": lambda: gemm_case(512, 320, 256, 256, rowtab=True, row_mode=True),
    "gemm_f32_cls": lambda: gemm_case(777, 512, 64, 64, bias=True, out_f32=True, nout=60),
    "gemm_many_tiles": lambda: gemm_case(148 * 128 * 3 + 77, 256, 768, 256),
    "gemm_big_time": lambda: gemm_case(98304, 256, 512, 256, act=1, bias=True, time_it=True),
    "gemm_big_time_row": lambda: gemm_case(98304, 512, 256, 256, bias=True, resid=True, ln1=True, row_mode=True, time_it=True),
    "gemm_big_time_row256": lambda: gemm_case(98304, 256, 256, 256, bias=True, resid=True, ln0=True, ln1=True, row_mode=True, time_it=True),
    "gemm_big_glu": lambda: gemm_case(98304, 256, 512, 256, act=3, bias=True, time_it=True),
    "gemm_scan_m": lambda: all([gemm_case(m, 256, 768, 256, time_it=True) for m in (18944, 37888, 75776, 151552, 303104)]),
    "gemm_big_qkv": lambda: gemm_case(98304, 256, 768, 256, time_it=True),
    "dw_eca": lambda: dw_case(3, 384, 512, 11, 10, 2),
    "dw_eca_k3": lambda: dw_case(2, 100, 128, 3, 2, 2),
    "dw_swish_colsum": lambda: dw_case(3, 384, 512, 15, 14, 1, bias=False, colsum=True),
    "dw_same": lambda: dw_case(2, 384, 256, 15, 7, 0),
    "attn": lambda: attn_case(2, 384, 8, 32),
    "attn_t256_mask": lambda: attn_case(3, 256, 4, 32, mask=True),
    "attn_big_time": lambda: attn_case(256, 384, 8, 32, time_it=True),
    "attn_mask_ragged": lambda: attn_case(2, 200, 4, 32, mask=True),
    "attn_dh48": lambda: attn_case(1, 1024, 8, 48),
    "ctc": lambda: ctc_case(6, 384, 60, 64),
    "ctc_small": lambda: ctc_case(3, 20, 8, 5),
    "decode": lambda: decode_case(5, 384, 60),
}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        results = {}
        for name in CASES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=180)
                out = (r.stdout + r.stderr).strip()
                status = "ok" if r.returncode == 0 else f"rc={r.returncode}"
            except subprocess.TimeoutExpired as e:
                out = ((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")) + " TIMEOUT"
                status = "timeout"
            results[name] = status
            print(f"===== {name} [{status}] {time.time()-t0:.1f}s\n{out[-3000:]}", flush=True)
        print("SUMMARY", results)
        sys.exit(0 if all(v == "ok" for v in results.values()) else 1)
    ok = CASES[which]()
    sys.exit(0 if ok else 1)
