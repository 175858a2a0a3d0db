"""Summarise DHG_TRACE output of the GEMM kernel (timeline of CTA 0).  python tools/trace_summary.py file [first last]"""
import sys
names = {0x10: 'P_slot', 0x20: 'M_tmem_ok', 0x21: 'M_a_ok', 0x22: 'M_commit', 0x30: 'E_wait', 0x31: 'E_full_ok', 0x32: 'E_done',
         0x33: ' e_chunk', 0x34: ' e_ld_done', 0x35: ' e_math_done', 0x36: ' e_stored'}
ev = []
for l in open(sys.argv[1]):
    if l.startswith('tc_gemm plan'): print(l.strip())
    if l.startswith('TR '):
        f = l.split(); t, c, i = f[1:4]; ev.append((int(t), int(c, 16), int(i)))
ev.sort()
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (6, 8)
print(len(ev), 'events; span', ev[-1][0] - ev[0][0], 'cycles')
show = False
for t, c, i in ev:
    if c == 0x10: continue
    if c == 0x30: show = lo <= i <= hi
    if c >= 0x30:
        if show: print(f"   {t:8d} {names[c]:12s} {i}")
    elif lo <= i <= hi: print(f"   {t:8d} {names[c]:12s} tile {i}")
ed = [t for t, c, i in ev if c == 0x32]
ew = [t for t, c, i in ev if c == 0x30]; ef = [t for t, c, i in ev if c == 0x31]
mt = [t for t, c, i in ev if c == 0x20]; mc = [t for t, c, i in ev if c == 0x22]
print('   tile period', (ed[-1] - ed[0]) / (len(ed) - 1))
print('   epilogue: wait for accumulator', sum(b - a for a, b in zip(ew, ef)) / len(ew), ' work', sum(d - f for d, f in zip(ed, ef)) / len(ed))
print('   MMA warp: tmem_ok -> commit', sum(c - t for t, c in zip(mt, mc)) / len(mt), ' commit -> next tmem_ok', sum(t - c for c, t in zip(mc, mt[1:])) / (len(mt) - 1))

# per-chunk phases of the first epilogue warp (codes 0x33..0x36)
ph = {c: [t for t, cc, i in ev if cc == c] for c in (0x33, 0x34, 0x35, 0x36)}
if ph[0x33] and len(ph[0x33]) == len(ph[0x34]) == len(ph[0x35]):
    n = len(ph[0x33])
    ld = sum(b - a for a, b in zip(ph[0x33], ph[0x34])) / n
    math = sum(b - a for a, b in zip(ph[0x34], ph[0x35])) / n
    st = sum(b - a for a, b in zip(ph[0x35], ph[0x36])) / max(len(ph[0x36]), 1) if len(ph[0x36]) == n else float('nan')
    print(f'   per chunk (warp 0 of the epilogue, {n} chunks): tcgen05.ld {ld:.0f}  bias/aux/FiLM {math:.0f}  pack + store {st:.0f} cycles')
