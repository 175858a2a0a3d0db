"""Timeline of one attention launch (CTA 0, slot 0) from the kernel's DHG_TRACE points.

python tools/attn_trace.py H Tq Tk self(0/1) masked(0/1) [B] [D]

Prints, averaged over the slot's items (first and last dropped), the time between consecutive events of the control
warp and of the first softmax warp, the item period of the slot and the launch time without tracing."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diffusion-handwriting-generation.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

NAMES = {0x02: "tiles landed", 0x03: "TMEM free", 0x05: "P ready (bar_p seen)", 0x07: "PV done (bar_o seen), next load issued",
         0x11: "item start", 0x12: "S ready (bar_s seen)", 0x13: "max pass done", 0x14: "exp pass done, P stored",
         0x15: "O ready (bar_o seen)", 0x16: "O stored, slot freed"}


def main():
    H, Tq, Tk, sa, mk = (int(x) for x in sys.argv[1:6])
    B = int(sys.argv[6]) if len(sys.argv) > 6 else 1024
    D = int(sys.argv[7]) if len(sys.argv) > 7 else 64
    from test_gpu_attention import run_attention
    from dhg_b200 import _abi
    lib = _abi.lib()
    if os.environ.get("ATTN_HALVES"):   # softmax warps per TMEM lane quarter
        lib.dhg_set_option(None, b"attn_dbg", -int(os.environ["ATTN_HALVES"]))
    _, _, ms = run_attention(lib, B, H, D, Tq, Tk, bool(sa), bool(mk), 1, seed=1, repeats=5)
    tmp = tempfile.NamedTemporaryFile(delete=False)
    saved = os.dup(2)
    os.dup2(tmp.fileno(), 2)
    os.environ["DHG_TRACE"] = "1"
    try:
        run_attention(lib, B, H, D, Tq, Tk, bool(sa), bool(mk), 1, seed=1, repeats=0)
    finally:
        os.dup2(saved, 2)
        del os.environ["DHG_TRACE"]
    ev = []
    slots = None
    for ln in open(tmp.name):
        f = ln.split()
        if ln.startswith("attention plan"):
            slots = ln.strip()
        if f and f[0] == "ATR":
            ev.append((int(f[1]), int(f[2], 16), int(f[3])))
    os.unlink(tmp.name)
    print(f"H={H} Tq={Tq} Tk={Tk} self={sa} masked={mk} B={B} D={D}: {ms * 1e3:.1f} us per launch; {slots}")
    for role, codes in ((0, (0x02, 0x03, 0x05, 0x07)), (1, (0x11, 0x12, 0x13, 0x14, 0x15, 0x16))):
        evs = sorted(e for e in ev if e[1] in codes)
        items = sorted({e[2] for e in evs})
        if len(items) < 4:
            print("  too few items traced")
            continue
        keep = set(items[1:-1])
        acc, cnt = {}, {}
        prev = None
        for t, c, it in evs:
            if prev is not None and it in keep:
                key = (prev[1], c)
                acc[key] = acc.get(key, 0) + t - prev[0]
                cnt[key] = cnt.get(key, 0) + 1
            prev = (t, c, it)
        starts = [t for t, c, it in evs if c == codes[0]]
        period = (starts[-1] - starts[0]) / (len(starts) - 1)
        print(f"  {'control warp' if role == 0 else 'softmax warp 0'}: {len(items)} items, period {period:.0f} cycles")
        for (a, b), v in acc.items():
            print(f"    {NAMES[a]:42s} -> {NAMES[b]:42s} {v / cnt[(a, b)]:8.0f}")


if __name__ == "__main__":
    main()
