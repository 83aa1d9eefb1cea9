"""Wallet-side construction of BlockMaze transactions through the library's own cgo helpers (genCMT, computePRF, genCMTS,
computeCRH) -- what go-ethereum/zktx does before it calls gen<Circuit>proof (zktx.go:122-260): serial numbers are PRF(sk, r),
note commitments are genCMT / genCMTS, the transfer randomness is CRH(pk_sender, r).  `synthetic` is the seeded generator of valid
transactions behind bench.py's workloads: random 32/20-byte secrets, value_old in [2^32, 2^40), value_s in
[1, value_old).  It draws exactly like tests/golden/fixtures.py (which hashes with the oracle instead), so both bench arms and the
parity tests see the same transactions; tests/test_host.py checks that the two generators agree."""
import random

from . import api


def _blob(s, nbytes=32):
    """A hex string the way zktx.go passes blobs: '0x' + 2*nbytes digits (uint256S semantics: short strings are zero-extended)."""
    s = s.strip().lower()
    if s.startswith("0x"):
        s = s[2:]
    return "0x" + s.rjust(2 * nbytes, "0")[-2 * nbytes:]


def _h(name, *args):
    return "0x" + api.helper(name, *args)


def mint(value, value_old, value_s, sk, r_old, r):
    """Arguments of genMintproof / genRedeemproof (SRC/mint/mintcgo.hpp:9-20)."""
    sk, r_old, r = _blob(sk), _blob(r_old), _blob(r)
    sn_old, sn = _h("computePRF", sk, r_old), _h("computePRF", sk, r)
    cmt_old, cmt = _h("genCMT", value_old, sn_old, r_old), _h("genCMT", value, sn, r)
    return [value, value_old, sn_old, r_old, sn, r, cmt_old, cmt, value_s, sk]


def send(value, value_old, value_s, sk, r_old, r, pk_sender, pk_recv):
    """Arguments of genSendproof (SRC/send/sendcgo.hpp:9-24); the 'A' note is the old one."""
    sk, r_old, r = _blob(sk), _blob(r_old), _blob(r)
    pk_sender, pk_recv = _blob(pk_sender, 20), _blob(pk_recv, 20)
    sn_old, sn = _h("computePRF", sk, r_old), _h("computePRF", sk, r)
    cmt_old, cmt = _h("genCMT", value_old, sn_old, r_old), _h("genCMT", value, sn, r)
    r_s = _h("computeCRH", pk_sender, r)
    cmt_s = _h("genCMTS", value_s, pk_recv, r_s, sn_old)
    return [value_old, r_s, sn_old, r_old, cmt_s, cmt_old, value_s, pk_recv, value, sn, r, cmt, sk, pk_sender]


def deposit(value, value_old, value_s, sk, r_old, r, r_s, pk_recv, sn_a_old, leaves, index):
    """Arguments of genDepositproof (SRC/deposit/depositcgo.hpp:9-28); leaves: '0x..' strings, leaves[index] is replaced by cmtS."""
    sk, r_old, r, r_s = _blob(sk), _blob(r_old), _blob(r), _blob(r_s)
    pk_recv, sn_a_old = _blob(pk_recv, 20), _blob(sn_a_old)
    sn_old, sn, sn_s = _h("computePRF", sk, r_old), _h("computePRF", sk, r), _h("computePRF", sk, r_s)
    cmt_old, cmt = _h("genCMT", value_old, sn_old, r_old), _h("genCMT", value, sn, r)
    cmt_s = _h("genCMTS", value_s, pk_recv, r_s, sn_a_old)
    leaves = list(leaves)
    leaves[index] = cmt_s
    rt = "0x" + "00" * 32                       # genDepositproof ignores RT (depositcgo.cpp:402-403)
    return [value, value_old, sn_old, r_old, sn, r, sn_s, r_s, cmt_old, cmt, value_s, pk_recv, sn_a_old, cmt_s, "".join(leaves), len(leaves), rt, sk]


def synthetic(circuit, seed):
    """Seeded valid transaction of the given type as a gen<Circuit>proof argument list."""
    rng = random.Random((seed << 2) | {"mint": 0, "send": 1, "deposit": 2, "redeem": 3}[circuit])
    hx = lambda n: "%0*x" % (2 * n, rng.getrandbits(8 * n))
    value_old = rng.randrange(1 << 32, 1 << 40)
    value_s = rng.randrange(1, value_old)
    kw = dict(sk=hx(32), r_old=hx(32), r=hx(32))
    if circuit == "mint":
        return mint(value_old + value_s, value_old, value_s, **kw)
    if circuit == "redeem":
        return mint(value_old - value_s, value_old, value_s, **kw)
    if circuit == "send":
        return send(value_old - value_s, value_old, value_s, pk_sender=hx(20), pk_recv=hx(20), **kw)
    n = 256
    leaves = ["0x" + bytes(reversed(bytes(rng.getrandbits(8) for _ in range(32)))).hex() for _ in range(n)]
    return deposit(value_old + value_s, value_old, value_s, r_s=hx(32), pk_recv=hx(20), sn_a_old=hx(32), leaves=leaves,
                   index=rng.randrange(n), **kw)
