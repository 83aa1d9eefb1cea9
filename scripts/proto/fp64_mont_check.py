"""Checks the output of fp64_mont against Python integers: r == a * b * 2^-260 mod q, fully reduced."""
import sys
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
RINV = pow(1 << 260, -1, Q)
n = bad = 0
for line in sys.stdin:
    a, b, r = (int(x, 16) for x in line.split())
    n += 1
    if r != a * b * RINV % Q:
        bad += 1
        if bad < 5:
            print("MISMATCH", hex(a), hex(b), hex(r), hex(a * b * RINV % Q))
print("%d products checked, %d mismatches" % (n, bad))
sys.exit(1 if bad else 0)
