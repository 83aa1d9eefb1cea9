"""The drop-in boundary seen from C: tests/cabi_harness/zktx_harness.c is compiled against the reference's own cgo headers
(go-ethereum/zktx/{mint,send,deposit,redeem}cgo.hpp, taken from /root/reference where they lie) and linked with the link line of
go-ethereum/zktx/zktx.go:4 against symlinks libzk_<c>.so -> libzkb200.so (INTEGRATION.md option A).  A signature drift is a compile /
link error here; Python's ctypes would not notice one.  The binary is built in the build container (__graft_entry__.build) and
travels to the GPU box, where /root/reference does not exist."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EXE = os.path.join(ROOT, "tests", "cabi_harness", "_build", "zktx_harness")


def _key_dir():
    for d in (os.environ.get("ZKB200_KEY_DIR"), os.path.join(ROOT, "oracle", "_ref", "prfKey"), "/usr/local/prfKey"):
        if d and os.path.exists(os.path.join(d, "mintvk.txt")):
            return d
    return None


def _run(args, tmp_path, timeout):
    import __graft_entry__ as G
    libs = G.link_dir(str(tmp_path / "lib"))                       # fresh symlinks + stubs (the snapshot need not preserve links)
    env = dict(os.environ, LD_LIBRARY_PATH=libs + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), ZKB200_KEY_DIR=_key_dir() or "")
    return subprocess.run([EXE] + args, env=env, capture_output=True, text=True, timeout=timeout)


def test_c_harness_compiles_against_reference_headers_and_links():
    """Build-container half: compile + link (only where /root/reference exists), then helpers and verifyMintproof -- host code, no GPU."""
    import __graft_entry__ as G
    if os.path.isdir("/root/reference/go-ethereum/zktx"):
        if os.path.exists(EXE):
            os.remove(EXE)
        assert G.build_cabi_harness() == EXE
    if not os.path.exists(EXE):
        pytest.skip("harness binary not built (needs /root/reference at build time)")
    needed = subprocess.run(["readelf", "-d", EXE], capture_output=True, text=True).stdout
    assert "libzk_mint.so" in needed and "libzkb200" not in needed           # bound by the reference's library names only


def test_c_harness_helpers_and_verify_on_host(tmp_path):
    from oracle import bn254_oracle as O
    if not os.path.exists(EXE) or _key_dir() is None:
        pytest.skip("harness binary / verification keys not present")
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "mint.json")))
    a = g["args"]
    out = _run(["host", g["proof_hex"], a[6], a[2], a[7], str(a[8])], tmp_path, 120)
    assert out.returncode == 0, out.stdout + out.stderr
    got = dict(l.split(" ", 1) for l in out.stdout.splitlines() if l.split(" ", 1)[0] in ("PRF", "CMT", "CRH", "CMTS", "ROOT", "VERIFY"))
    sk, r = bytes(31) + b"\x01", bytes([0x12] * 31 + [0x34])
    u = lambda b: O.uint256S(b.hex())                           # the in-memory blob of a "0x..." argument
    sn = O.compute_prf(u(sk), u(r))
    assert got["PRF"] == O.blob_hex(sn)
    cmt = O.note_cm(13, sn, u(r))
    assert got["CMT"] == O.blob_hex(cmt)
    pk7, pk9 = bytes((7 + i) & 0xff for i in range(20)), bytes((9 + i) & 0xff for i in range(20))
    crh = O.compute_crh(O.uint256S(pk7.hex(), 20), u(r))
    assert got["CRH"] == O.blob_hex(crh)
    cmts = O.notes_cm(5, O.uint256S(pk9.hex(), 20), crh, sn)
    assert got["CMTS"] == O.blob_hex(cmts)
    level = [cmt, cmts] + [bytes(32)] * 254
    while len(level) > 1:
        level = [O.sha256_compress(level[i], level[i + 1]) for i in range(0, len(level), 2)]
    assert got["ROOT"] == O.blob_hex(level[0])
    assert got["VERIFY"] == "1 0"


@pytest.mark.gpu
def test_c_harness_full_round_trip_on_gpu(tmp_path):
    """computePRF -> genCMT -> gen<X>proof -> verify<X>proof for the four circuits, called from C exactly as zktx.go encodes them."""
    assert os.path.exists(EXE), "tests/cabi_harness/_build/zktx_harness must travel to the GPU box (built by __graft_entry__.build)"
    assert _key_dir() is not None
    out = _run(["full"], tmp_path, 600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    for tag in ("MINT 1", "MINT_UNSAT 1", "REDEEM 1", "SEND 1", "DEPOSIT 1"):
        assert tag in out.stdout, out.stdout[-2000:]
