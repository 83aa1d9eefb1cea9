"""The reference's own end-to-end fixtures (SRC/<circuit>/main.cpp) expressed as gen<Circuit>proof argument lists,
plus a seeded generator of valid synthetic transactions.  Uses only hashlib-level helpers of
the oracle (uint256S / Note::cm / PRF / CRH restatements)."""
import random
from oracle import bn254_oracle as O

H = O.arg_hex


def mint_fixture(value=13, value_old=6, value_s=7, sk="1", r_old="123456", r="123"):
    """SRC/mint/main.cpp:121-132,249-251."""
    sk, r_old, r = O.uint256S(sk), O.uint256S(r_old), O.uint256S(r)
    sn_old, sn = O.compute_prf(sk, r_old), O.compute_prf(sk, r)
    cmtA_old, cmtA = O.note_cm(value_old, sn_old, r_old), O.note_cm(value, sn, r)
    return [value, value_old, H(sn_old), H(r_old), H(sn), H(r), H(cmtA_old), H(cmtA), value_s, H(sk)]


def redeem_fixture(value=13, value_old=20, value_s=7, sk="1", r_old="123456", r="123"):
    """SRC/redeem/main.cpp (same shape as mint; value = value_old - value_s), values :249-251."""
    return mint_fixture(value, value_old, value_s, sk, r_old, r)


def send_fixture(value=14, value_old=22, value_s=8, sk="1", r_old="123456", r="12", pk_sender="456", pk_recv="123"):
    """SRC/send/main.cpp:113-131,261-263.  genSendproof(value_A, r_s, sn, r, cmt_s, cmtA, value_s, pk_recv, value_A_new,
    sn_A_new, r_A_new, cmt_A_new, sk, pk_sender): the 'A' note is the OLD note, 'A_new' the new one."""
    sk, r_old, r = O.uint256S(sk), O.uint256S(r_old), O.uint256S(r)
    pk_sender, pk_recv = O.uint256S(pk_sender, 20), O.uint256S(pk_recv, 20)
    sn_old, sn = O.compute_prf(sk, r_old), O.compute_prf(sk, r)
    cmtA_old, cmtA = O.note_cm(value_old, sn_old, r_old), O.note_cm(value, sn, r)
    r_s = O.compute_crh(pk_sender, r)
    cmtS = O.notes_cm(value_s, pk_recv, r_s, sn_old)
    return [value_old, H(r_s), H(sn_old), H(r_old), H(cmtS), H(cmtA_old), value_s, H(pk_recv), value, H(sn), H(r), H(cmtA),
            H(sk), H(pk_sender)]


def deposit_fixture(value=264, value_old=255, value_s=9, sk="1", r_old="123456", r="12", r_s="123", pk_recv="123",
                    sn_A_old="123", leaves=None, index=9):
    """SRC/deposit/main.cpp:120-185,331-333: 16 leaves "1".."16" with cmtS at index 9."""
    sk, r_old, r, r_s = O.uint256S(sk), O.uint256S(r_old), O.uint256S(r), O.uint256S(r_s)
    pk_recv, sn_A_old = O.uint256S(pk_recv, 20), O.uint256S(sn_A_old)
    sn_old, sn, sn_s = O.compute_prf(sk, r_old), O.compute_prf(sk, r), O.compute_prf(sk, r_s)
    cmtB_old, cmtB = O.note_cm(value_old, sn_old, r_old), O.note_cm(value, sn, r)
    cmtS = O.notes_cm(value_s, pk_recv, r_s, sn_A_old)
    if leaves is None:
        leaves = [O.uint256S(str(i + 1)) for i in range(16)]
    leaves = list(leaves)
    leaves[index] = cmtS
    cmtarray = "".join(H(x) for x in leaves)
    rt = H(bytes(32))     # genDepositproof ignores RT (depositcgo.cpp:402-403)
    return [value, value_old, H(sn_old), H(r_old), H(sn), H(r), H(sn_s), H(r_s), H(cmtB_old), H(cmtB), value_s, H(pk_recv),
            H(sn_A_old), H(cmtS), cmtarray, len(leaves), rt, H(sk)]


FIXTURES = {"mint": mint_fixture, "send": send_fixture, "deposit": deposit_fixture, "redeem": redeem_fixture}


def synthetic(circuit, seed):
    """Seeded valid transaction of the given type: random 32/20-byte secrets, value_old in [2^32, 2^40),
    value_s in [1, value_old)."""
    rng = random.Random((seed << 2) | {"mint": 0, "send": 1, "deposit": 2, "redeem": 3}[circuit])
    hx = lambda n: "%0*x" % (2 * n, rng.getrandbits(8 * n))
    value_old = rng.randrange(1 << 32, 1 << 40)
    value_s = rng.randrange(1, value_old)
    kw = dict(sk=hx(32), r_old=hx(32), r=hx(32))
    if circuit == "mint":
        return mint_fixture(value_old + value_s, value_old, value_s, **kw)
    if circuit == "redeem":
        return mint_fixture(value_old - value_s, value_old, value_s, **kw)
    if circuit == "send":
        return send_fixture(value_old - value_s, value_old, value_s, pk_sender=hx(20), pk_recv=hx(20), **kw)
    n = 256
    leaves = [bytes(rng.getrandbits(8) for _ in range(32)) for _ in range(n)]
    return deposit_fixture(value_old + value_s, value_old, value_s, r_s=hx(32), pk_recv=hx(20), sn_A_old=hx(32),
                           leaves=leaves, index=rng.randrange(n), **kw)
