"""Pins the CPU restatement (oracle/bn254_oracle.py) against the UNMODIFIED reference (oracle/_ref, skipped where the
reference build is absent) and against the committed golden vectors generated from the reference."""
import hashlib
import json
import os
import random
import zlib

import pytest

from oracle import bn254_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CIRCUITS = ("mint", "send", "deposit", "redeem")


def gold(c):
    return json.load(open(os.path.join(GOLD, c + ".json")))


def test_golden_randomness_restatement():
    """Fr::random_element word consumption (fp.tcc:695-721): the pinned word stream must give the golden (r, s)."""
    for c in CIRCUITS:
        g = gold(c)
        assert g["words"] == O.fixed_rng_words(g["seed"], len(g["words"]))
        r, s = O.fr_from_words(g["words"])
        assert (r, s) == (int(g["r"], 16), int(g["s"], 16))


def test_golden_assignments_match_digest_and_sizes():
    sizes = {"mint": (151512, 196608), "send": (227046, 262144), "deposit": (457127, 524288), "redeem": (151579, 196608)}
    for c in CIRCUITS:
        g = gold(c)
        w = zlib.decompress(open(os.path.join(GOLD, c + "_assignment.bin.z"), "rb").read())
        assert hashlib.sha256(w).hexdigest() == g["assignment_sha256"]
        assert (len(w) // 32, g["domain_size"]) == sizes[c]
        assert g.get("cgo_genproof_equal", True)


def test_golden_proof_points_on_curve():
    for c in CIRCUITS:
        h = gold(c)["proof_hex"]
        v = [int(h[64 * i:64 * i + 64], 16) for i in range(8)]
        assert O.G1.on_curve((v[0], v[1])) and O.G1.on_curve((v[6], v[7]))
        assert O.G2.on_curve((O.Fq2(v[3], v[2]), O.Fq2(v[5], v[4])))


def test_helper_known_answers():
    """values captured from the reference's computePRF / genCMT."""
    sk, r_old, r = O.uint256S("1"), O.uint256S("123456"), O.uint256S("123")
    sn_old, sn = O.compute_prf(sk, r_old), O.compute_prf(sk, r)
    assert O.blob_hex(sn_old) == "4a31770fe5354a1a9632ebe1481e108cd82ce514ac094c57b5ffdfaea8ac138a"
    assert O.blob_hex(sn) == "59416ca7b4d0fdcb61dd7fb063db35e9a0a96dd9fecf20a8027aed2d6c4f4006"
    assert O.blob_hex(O.note_cm(6, sn_old, r_old)) == "53996012d011396f7a4953c80ee64d5a1b8b7ed676bbea7886b698c09aa991db"
    assert O.blob_hex(O.note_cm(13, sn, r)) == "dbc961ea0d748198f21ea3267534e064eaaea65733dbd362fed93b4748991c27"


def test_domain_selection_rule():
    """get_evaluation_domain.tcc:33-52 for the four circuits and the rounding cases."""
    for ms, m, kind in ((167275, 196608, "step_radix2"), (252292, 262144, "basic_radix2"), (503870, 524288, "basic_radix2"),
                        (167858, 196608, "step_radix2"), (8, 8, "basic_radix2"), (12, 12, "step_radix2"), (100, 128, "basic_radix2"),
                        (3000, 3072, "step_radix2"), (5000, 5120, "step_radix2")):
        d = O.get_evaluation_domain(ms)
        assert (d.m, d.kind) == (m, kind)


@pytest.mark.parametrize("ms", [2, 8, 12, 24, 64, 768, 1000, 3000])
def test_domain_properties(ms):
    """The libfqfft gtest properties (tests/evaluation_domain_test.cpp:40-236): FFT == naive evaluation on the domain,
    iFFT(FFT) = id, icosetFFT(cosetFFT) = id, Z vanishes on the domain."""
    rng = random.Random(ms)
    d = O.get_evaluation_domain(ms)
    a = [rng.randrange(O.R_MOD) for _ in range(d.m)]
    ev = d.FFT(a)
    for i in random.Random(1).sample(range(d.m), min(d.m, 6)):
        x = d.get_domain_element(i)
        assert ev[i] == sum(c * pow(x, k, O.R_MOD) for k, c in enumerate(a)) % O.R_MOD
        assert d.compute_vanishing_polynomial(x) == 0
    assert d.iFFT(ev) == a
    assert d.icosetFFT(d.cosetFFT(a, 5), 5) == a


def _random_r1cs(rng, num_inputs, num_aux, nc):
    """A satisfied random R1CS: each row (a.w)*(b.w) = c.w with c's last coefficient solved for a fresh aux variable."""
    n = num_inputs + num_aux
    full = [rng.randrange(O.R_MOD) for _ in range(n)]
    rows = []
    for i in range(nc):
        def lc():
            return [(rng.randrange(0, n + 1), rng.choice([1, O.R_MOD - 1, rng.randrange(O.R_MOD)])) for _ in range(rng.randrange(1, 4))]
        a, b = lc(), lc()
        prod = O.lc_eval(a, full) * O.lc_eval(b, full) % O.R_MOD
        # c = k * w_j with k chosen so the row holds
        j = rng.randrange(1, n + 1)
        wj = full[j - 1] or 1
        full[j - 1] = wj
        rows.append((a, b, [(j, prod * pow(wj, -1, O.R_MOD) % O.R_MOD)]))
    # re-solve rows whose inputs were touched later: simply recompute c coefficients in order
    fixed = []
    for a, b, c in rows:
        j = c[0][0]
        prod = O.lc_eval(a, full) * O.lc_eval(b, full) % O.R_MOD
        fixed.append((a, b, [(j, prod * pow(full[j - 1], -1, O.R_MOD) % O.R_MOD)]))
    return O.R1CS(num_inputs, num_aux, fixed), full


@pytest.mark.parametrize("shape", [(3, 20, 27), (4, 40, 59), (2, 30, 93)])
def test_qap_witness_map_against_reference(ref, shape):
    """r1cs_to_qap_witness_map (r1cs_to_qap.tcc:205-334) on small random systems, basic and step domains."""
    rng = random.Random(sum(shape))
    cs, full = _random_r1cs(rng, *shape)
    assert O.is_satisfied(cs, full)
    mine = O.r1cs_to_qap_witness_map(cs, full)
    theirs, _ = ref.qap_witness_map(cs, full)
    assert mine == theirs
    assert mine[-1] == 0 and mine[-2] == 0     # degree(H) = m - 2 (r1cs_gg_ppzksnark.tcc:406-408)


def test_domains_against_reference(ref):
    rng = random.Random(5)
    for ms in (8, 12, 100, 768, 3000):
        m, kind = ref.domain_size(ms)
        d = O.get_evaluation_domain(ms)
        assert (d.m, d.kind) == (m, kind)
        v = [rng.randrange(O.R_MOD) for _ in range(m)]
        assert ref.domain_op(ms, "FFT", v) == d.FFT(v)
        assert ref.domain_op(ms, "iFFT", v) == d.iFFT(v)
        assert ref.domain_op(ms, "cosetFFT", v) == d.cosetFFT(v, 5)
        assert ref.domain_op(ms, "icosetFFT", v) == d.icosetFFT(v, 5)
        assert ref.domain_op(ms, "divide_by_Z_on_coset", v) == d.divide_by_Z_on_coset(v)


def test_montgomery_representation_against_reference(ref):
    rng = random.Random(9)
    a = [0, 1, O.R_MOD - 1] + [rng.randrange(O.R_MOD) for _ in range(8)]
    assert ref.to_mont("fr", a) == [O.fr_to_mont(x) for x in a]
    b = [0, 1, O.Q_MOD - 1] + [rng.randrange(O.Q_MOD) for _ in range(8)]
    assert ref.to_mont("fq", b) == [O.fq_to_mont(x) for x in b]


def test_msm_against_reference(ref):
    rng = random.Random(11)
    n = 200
    b1, b2 = ref.g1_bases_bytes(n, 12345), ref.g2_bases_bytes(n, 999)
    sc = [rng.randrange(O.R_MOD) for _ in range(n)]
    for i in range(0, n, 3):
        sc[i] = rng.choice([0, 1, rng.randrange(1 << 33)])
    p1 = [ref.g1_from(b1[64 * i:64 * i + 64]) for i in range(n)]
    p2 = [ref.g2_from(b2[128 * i:128 * i + 128]) for i in range(n)]
    assert all(O.G1.on_curve(p) for p in p1) and all(O.G2.on_curve(p) for p in p2)
    out, _ = ref.msm_g1_bytes(b1, ref.fr_bytes(sc), 1)
    assert O.G1.to_affine(O.G1.multi_exp_with_mixed_addition([O.G1.from_affine(p) for p in p1], sc)) == ref.g1_from(out)
    out, _ = ref.msm_g2_bytes(b2, ref.fr_bytes(sc), 0)
    assert O.G2.to_affine(O.G2.multi_exp_inner([O.G2.from_affine(p) for p in p2], sc)) == ref.g2_from(out)


def test_pk_header_and_prover_pieces_against_golden(ref):
    """Parse the head of the reference mint pk with the oracle's parser; the first A/L points must be on the curve and the
    constraint system must be satisfied by the golden assignment (oracle is_satisfied restatement, r1cs.tcc:133-164)."""
    path = os.path.join(ref.KEY_DIR, "mintpk.txt")
    if not os.path.exists(path):
        pytest.skip("reference keys not present")
    pk = O.parse_pk(open(path, "rb").read(), max_points=4)
    assert pk.cs.num_variables == 151512 and pk.cs.num_inputs == 4 and pk.cs.num_constraints == 167270
    assert all(O.G1.on_curve(p) for p in pk.A_query + pk.H_query + pk.L_query + pk.B_g1)
    assert all(O.G2.on_curve(p) for p in pk.B_g2)
    w = zlib.decompress(open(os.path.join(GOLD, "mint_assignment.bin.z"), "rb").read())
    full = [int.from_bytes(w[32 * i:32 * i + 32], "little") for i in range(len(w) // 32)]
    rng = random.Random(3)
    rows = rng.sample(pk.cs.rows, 3000)
    assert all(O.lc_eval(a, full) * O.lc_eval(b, full) % O.R_MOD == O.lc_eval(c, full) for a, b, c in rows)


def test_sha512_rng_restatement_equals_libff(ref):
    """The scalar stream of the synthetic kernel sweep: oracle.sha512_rng == libff::SHA512_rng<Fr> (rng.tcc:26-72), incl. indices whose first
    digests are rejected (a quarter of all draws) and the 2^32 + i range the bases use."""
    for first in (0, 12345, 1 << 32, (1 << 40) + 3):
        assert ref.sha512_rng(first, 300) == [O.sha512_rng(first + i) for i in range(300)]
