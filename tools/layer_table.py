"""Per-layer table of the U-Net's conv launches from an `ncu --metrics gpu__time_duration.sum` launch
list of bench.py (one forward, batch 8): layer shape -> kernel variant, time, TFLOP/s.
  python tools/layer_table.py profiles/r01k_launches.csv > profiles/r01k_layer_table.txt
The layer sequence is enumerated from the model definition (same order as the engine issues it:
stem, then per ResBlock in_layers.2 and out_layers.3 [+ skip_connection as extra K], per
AttentionBlock qkv and proj_out)."""
import collections
import csv
import sys

B, MC = 8, 256
MULT, NRB, ATTN_DS = (1, 1, 2, 2, 4, 4), 2, {8, 16, 32}


def layers():
    out = [("stem", 256, 64, MC, 1)]
    ch, ds, chans = MC, 1, [MC]
    H = 256

    def res(tag, cin, cout, up=False, down=False):
        nonlocal H
        if up:
            H *= 2
        if down:
            H //= 2
        out.append((tag + ".in2", H, cin, cout, 9))
        k = 9 * cout + (cin if cin != cout else 0)
        out.append((tag + (".out3+skip" if cin != cout else ".out3"), H, k / 9.0, cout, 9))

    def attn(tag, c):
        out.append((tag + ".qkv", H, c, 3 * c, 1))
        out.append((tag + ".proj", H, c, c, 1))

    for lvl, m in enumerate(MULT):
        for i in range(NRB):
            res(f"in{lvl}.{i}", ch, m * MC)
            ch = m * MC
            if ds in ATTN_DS:
                attn(f"in{lvl}.{i}", ch)
            chans.append(ch)
        if lvl != len(MULT) - 1:
            res(f"in{lvl}.down", ch, ch, down=True)
            chans.append(ch)
            ds *= 2
    res("mid.0", ch, ch)
    attn("mid", ch)
    res("mid.2", ch, ch)
    for lvl, m in list(enumerate(MULT))[::-1]:
        for i in range(NRB + 1):
            res(f"out{lvl}.{i}", ch + chans.pop(), m * MC)
            ch = m * MC
            if ds in ATTN_DS:
                attn(f"out{lvl}.{i}", ch)
            if lvl and i == NRB:
                res(f"out{lvl}.up", ch, ch, up=True)
                ds //= 2
    return out


rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 10 and r[0].isdigit()]
rows = [r for r in rows if r[-3] == "gpu__time_duration.sum"]  # lists with DRAM metrics hold three rows per launch
heads = [i for i, r in enumerate(rows) if "head_kernel" in r[4]]
# a bench.py launch list holds several forwards delimited by head_kernel launches; a conv-only
# list (ncu -k regex:conv..., one forward of tools/bench_unet.py) is taken whole
fw = rows[heads[0] + 1:heads[1] + 1] if len(heads) >= 2 else rows
launches = []
for r in fw:
    name = r[4].split("(")[0].replace("void ", "")
    us = float(r[-1]) / 1000
    if "splitk_finish" in name and launches:
        launches[-1][2] += us  # the finishing pass belongs to the split-K conv before it
    elif "conv_tc" in name or "conv_halo" in name:
        launches.append([name, r[8], us])
L = layers()
assert len(L) == len(launches), (len(L), len(launches))
agg = collections.OrderedDict()
for (tag, H, cin, cout, taps), (k, grid, us) in zip(L, launches):
    fl = 2.0 * B * H * H * cout * taps * cin
    key = (H, int(round(cin * taps)) if taps == 9 else cin, cout, taps, tag.split(".")[-1], k)
    a = agg.setdefault(key, [0, 0.0, fl])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(launches)} conv launches of one forward (batch 8), {tot / 1000:.3f} ms under ncu")
print(f"{'H':>4} {'K':>6} {'Cout':>5} taps {'layer':10s} {'kernel':32s} {'n':>2} {'avg_us':>8} {'TFLOP/s':>8} {'tot_us':>8}")
for (H, K, cout, taps, kind, k), (n, us, fl) in agg.items():
    print(f"{H:4d} {K:6d} {cout:5d} {taps:4d} {kind:10s} {k:32s} {n:2d} {us / n:8.1f} {fl / (us / n) / 1e6:8.0f} {us:8.1f}")
