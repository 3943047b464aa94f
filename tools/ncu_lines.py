"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
items, tot, tots = [], 0.0, 0.0
for r in rows[hi + 1:]:
    if len(r) <= iI or r[2] != "-":      # keep only the per-source-line summary rows (Address == '-')
        continue
    try:
        n, s = float(r[iI]), float(r[iS])
    except ValueError:
        continue
    tot += n; tots += s
    items.append((s, n, r[0], r[1][:100]))
items.sort(reverse=True)
print(f"total inst {tot:.3g}, samples {tots:.0f}")
for s, n, ln, src in items[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{s / tots * 100:5.1f}% samp {n / tot * 100:5.1f}% inst  L{ln}: {src}")
