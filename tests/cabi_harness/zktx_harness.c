/* TEST INFRASTRUCTURE -- a C caller standing in for go-ethereum/zktx/zktx.go (the Go toolchain is not in this image).
 *
 * It #includes the four headers cgo includes (go-ethereum/zktx/{mint,send,deposit,redeem}cgo.hpp, compiled from /root/reference where
 * they lie -- never copied), so a drift between libzkb200's exported signatures and the reference's shows up as a compile or link
 * error, which a ctypes binding cannot catch; it is linked with the reference's own cgo link line (zktx.go:4) against a directory in
 * which libzk_{mint,send,deposit,redeem}.so are symlinks to libzkb200.so (option A); and it encodes arguments exactly
 * as zktx.go does: 256-bit values as "0x" + 64 lowercase hex (common.ToHex, zktx.go:387-397), addresses as "0x" + 40 hex, results
 * taken with the equivalent of C.GoString.
 *
 *   zktx_harness host  <golden mint proof> <cmtA_old> <sn_old> <cmtA> <value_s>     helpers + verifyMintproof only (no GPU)
 *   zktx_harness full                                                               computePRF -> genCMT -> gen*proof -> verify*proof, 4 circuits
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mintcgo.hpp"
#include "sendcgo.hpp"
#include "depositcgo.hpp"
#include "redeemcgo.hpp"

/* common.ToHex of a 32-byte hash whose bytes are all `b` except the last, which is `last` */
static char *hash_hex(unsigned b, unsigned last) {
    char *s = malloc(67);
    strcpy(s, "0x");
    for (int i = 0; i < 31; i++) sprintf(s + 2 + 2 * i, "%02x", b & 0xff);
    sprintf(s + 64, "%02x", last & 0xff);
    return s;
}
static char *addr_hex(unsigned b) {
    char *s = malloc(43);
    strcpy(s, "0x");
    for (int i = 0; i < 20; i++) sprintf(s + 2 + 2 * i, "%02x", (b + i) & 0xff);
    return s;
}
/* zktx.go prefixes what the helpers return with "0x" before passing it on (common.ToHex of the parsed hash) */
static char *ox(const char *hex64) {
    char *s = malloc(67);
    strcpy(s, "0x");
    memcpy(s + 2, hex64, 64);
    s[66] = 0;
    return s;
}
static int is_default(const char *proof) { return strncmp(proof, "0000000000", 10) == 0; }      /* internal/ethapi/api.go:1486 */

static int run_host(int argc, char **argv) {
    if (argc < 7) return 2;
    char *sk = hash_hex(0, 1), *r = hash_hex(0x12, 0x34);
    char *sn = computePRF(sk, r);
    char *cmt = genCMT(13, ox(sn), r);
    char *crh = computeCRH(addr_hex(7), r);
    char *cmts = genCMTS(5, addr_hex(9), ox(crh), ox(sn));
    char leaves[2 * 66 + 1];
    snprintf(leaves, sizeof leaves, "%s%s", ox(cmt), ox(cmts));
    char *root = genRoot(leaves, 2);
    printf("PRF %.64s\nCMT %.64s\nCRH %.64s\nCMTS %.64s\nROOT %.64s\n", sn, cmt, crh, cmts, root);
    const bool ok = verifyMintproof(argv[2], argv[3], argv[4], argv[5], strtoull(argv[6], NULL, 10));
    char *tampered = strdup(argv[2]);
    tampered[300] = tampered[300] == '1' ? '2' : '1';
    const bool bad = verifyMintproof(tampered, argv[3], argv[4], argv[5], strtoull(argv[6], NULL, 10));
    printf("VERIFY %d %d\n", ok ? 1 : 0, bad ? 1 : 0);
    return ok && !bad ? 0 : 1;
}

static int run_full(void) {
    int fails = 0;
    char *sk = hash_hex(0xa1, 1), *r_old = hash_hex(0xb2, 2), *r_new = hash_hex(0xc3, 3);
    const uint64_t v_old = 1000, v_s = 77;
    char *sn_old = ox(computePRF(sk, r_old)), *sn_new = ox(computePRF(sk, r_new));
    char *cmt_old = ox(genCMT(v_old, sn_old, r_old));

    /* mint: value = value_old + value_s (zktx.go:383-404) */
    char *cmt_mint = ox(genCMT(v_old + v_s, sn_new, r_new));
    char *p = genMintproof(v_old + v_s, v_old, sn_old, r_old, sn_new, r_new, cmt_old, cmt_mint, v_s, sk);
    int ok = !is_default(p) && strlen(p) == 512 && verifyMintproof(p, cmt_old, sn_old, cmt_mint, v_s) && !verifyMintproof(p, cmt_old, sn_old, cmt_mint, v_s + 1);
    printf("MINT %d\n", ok); fails += !ok;
    p = genMintproof(v_old + v_s + 1, v_old, sn_old, r_old, sn_new, r_new, cmt_old, cmt_mint, v_s, sk);      /* values do not add up */
    ok = is_default(p);
    printf("MINT_UNSAT %d\n", ok); fails += !ok;

    /* redeem: value = value_old - value_s */
    char *cmt_red = ox(genCMT(v_old - v_s, sn_new, r_new));
    p = genRedeemproof(v_old - v_s, v_old, sn_old, r_old, sn_new, r_new, cmt_old, cmt_red, v_s, sk);
    ok = !is_default(p) && verifyRedeemproof(p, cmt_old, sn_old, cmt_red, v_s) && !verifyMintproof(p, cmt_old, sn_old, cmt_red, v_s);
    printf("REDEEM %d\n", ok); fails += !ok;

    /* send (zktx.go:406-440): r_s = CRH(pk_sender, r_new); cmtS = CMTS(value_s, pk_recv, r_s, sn_old) */
    char *pk_sender = addr_hex(0x40), *pk_recv = addr_hex(0x80);
    char *r_s = ox(computeCRH(pk_sender, r_new));
    char *cmt_s = ox(genCMTS(v_s, pk_recv, r_s, sn_old));
    p = genSendproof(v_old, r_s, sn_old, r_old, cmt_s, cmt_old, v_s, pk_recv, v_old - v_s, sn_new, r_new, cmt_red, sk, pk_sender);
    ok = !is_default(p) && verifySendproof(p, cmt_old, sn_old, cmt_s, cmt_red) && !verifySendproof(p, cmt_old, sn_old, cmt_red, cmt_s);
    printf("SEND %d\n", ok); fails += !ok;

    /* deposit (zktx.go:442-500): cmtS somewhere among the leaves; RT = genRoot(leaves) */
    enum { N = 9 };
    char *leaves = malloc(N * 66 + 1);
    leaves[0] = 0;
    for (int i = 0; i < N; i++) strcat(leaves, i == 5 ? cmt_s : hash_hex(0x30 + i, i));
    char *rt = ox(genRoot(leaves, N));
    char *r_dep = hash_hex(0xd4, 4);
    char *sn_s = ox(computePRF(sk, r_s)), *sn_dep = ox(computePRF(sk, r_dep));
    char *cmt_dep = ox(genCMT(v_old + v_s, sn_dep, r_dep));
    p = genDepositproof(v_old + v_s, v_old, sn_old, r_old, sn_dep, r_dep, sn_s, r_s, cmt_old, cmt_dep, v_s, pk_recv, sn_old, cmt_s, leaves, N, rt, sk);
    ok = !is_default(p) && verifyDepositproof(p, rt, pk_recv, cmt_old, sn_old, cmt_dep, sn_s) && !verifyDepositproof(p, rt, pk_sender, cmt_old, sn_old, cmt_dep, sn_s);
    printf("DEPOSIT %d\n", ok); fails += !ok;
    return fails;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "host") == 0) return run_host(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "full") == 0) return run_full();
    fprintf(stderr, "usage: zktx_harness host <proof> <cmtA_old> <sn_old> <cmtA> <value_s> | full\n");
    return 2;
}
