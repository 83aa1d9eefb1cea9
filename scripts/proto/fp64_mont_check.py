"""Checks the output of fp64_mont against Python integers: r == a * b * 2^-260 mod q, fully reduced.
   --lazy (scripts/ubench/mix host N): r < 2q and congruent, inputs below 2^256."""
import sys
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
RINV = pow(1 << 260, -1, Q)
LAZY = '--lazy' in sys.argv
n = bad = 0
for line in sys.stdin:
    a, b, r = (int(x, 16) for x in line.split())
    n += 1
    if (r >= 2 * Q or r % Q != a * b * RINV % Q) if LAZY else (r != a * b * RINV % Q):
        bad += 1
        if bad < 5:
            print("MISMATCH", hex(a), hex(b), hex(r), hex(a * b * RINV % Q))
print("%d products checked, %d mismatches" % (n, bad))
sys.exit(1 if bad else 0)
