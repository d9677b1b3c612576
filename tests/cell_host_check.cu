// tests/cell_host_check.cu -- the packed cell recipe of the product kernel (rabbitsalign_b200/csrc/fast_cell.cuh: the very
// function fast_dp_kernel calls) run on the HOST against a plain integer restatement of one Smith-Waterman cell of the
// reference (GASAL2/src/kernels/local_kernel_template.h:45-60: H, E', F', the 4-bit direction code), over random and
// boundary inputs for several scorings.  Checks the arithmetic the recipe rests on: signed profile bytes with the
// always-carry add, the always-borrow negations, clamp facts, HFMA2 merging of facts on denormal bit patterns, and the
// byte-permute gather of four columns.
//
//   nvcc -O1 -std=c++17 -o cell_host_check tests/cell_host_check.cu && ./cell_host_check      (exit 0 = all equal)
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../rabbitsalign_b200/csrc/fast_cell.cuh"

using namespace rsa;

static int half_s(uint32_t v, int h) { return (int)(int16_t)(h ? (v >> 16) : (v & 0xFFFFu)); }

struct Ref { int h, fn, en, nib; };
// one cell, true (unbiased) values; sub = substitution score of the cell
static Ref ref_cell(const Scoring& sc, int diag, int F, int e, int sub) {
    Ref r;
    const int tmp = diag + sub;
    const int tg = tmp - sc.gap_oe;
    int h = tmp > F ? tmp : F;
    h = h > e ? h : e;
    h = h > 0 ? h : 0;
    const int u = (F > e ? F : e) > 0 ? (F > e ? F : e) : 0;
    r.h = h;
    r.nib = 0;
    if (!(tg > F - sc.gap_ext)) r.nib |= 8;
    if (!(tg > e - sc.gap_ext)) r.nib |= 4;
    if (h != tmp) r.nib |= 2;         // H != diagonal  (kernel: max(F,E,0) > diag + sub)
    if (u != F) r.nib |= 1;           // max(F,E,0) != F
    r.fn = tg > F - sc.gap_ext ? tg : F - sc.gap_ext;
    r.en = tg > e - sc.gap_ext ? tg : e - sc.gap_ext;
    return r;
}

int main() {
    std::mt19937_64 rng(12345);
    long long checked = 0, bad = 0;
    const Scoring scorings[] = {{2, 8, 12, 1}, {1, 4, 7, 2}, {3, 5, 9, 0}, {5, 30, 30, 3}, {2, 8, 12, 12}};
    for (const Scoring& sc : scorings) {
        const FastConsts k = make_fast_consts(sc);
        const int floor_ef = -(sc.mismatch + sc.gap_oe);
        for (int it = 0; it < 400000; ++it) {
            // two independent cells (pair A, pair B) packed into one call
            int diag[2], F[2], e[2], tcode[2], qcode[2];
            bool qn[2];
            for (int p = 0; p < 2; ++p) {
                const int hi = (it & 7) == 0 ? 1023 : ((it & 7) == 1 ? 3 : 400);
                diag[p] = (int)(rng() % (hi + 1));
                F[p] = floor_ef + (int)(rng() % (hi + 1 - floor_ef));
                e[p] = floor_ef + (int)(rng() % (hi + 1 - floor_ef));
                if ((rng() & 3) == 0) F[p] = e[p];               // ties
                if ((rng() & 7) == 0) F[p] = diag[p] - sc.gap_oe + (int)(rng() % 5) - 2;
                if ((rng() & 7) == 0) e[p] = diag[p] - sc.gap_oe + (int)(rng() % 5) - 2;
                tcode[p] = (int)(rng() % 6);                     // 0..3 base, 4 N, 5 row past the window
                qcode[p] = (int)(rng() % 4);
                qn[p] = (rng() % 9) == 0;                        // N in the query
            }
            // substitution score as the reference defines it (gasal_kernels.h:48-56 without N_PENALTY)
            int sub[2];
            for (int p = 0; p < 2; ++p) {
                if (tcode[p] == 5) sub[p] = -sc.mismatch;
                else if (tcode[p] == 4 || qn[p]) sub[p] = 0;
                else sub[p] = (tcode[p] == qcode[p]) ? sc.match : -sc.mismatch;
                if (tcode[p] == 5 && qn[p]) sub[p] = 0;  // (the kernel's N mask wins; such rows never matter: see profile_word)
            }
            // the kernel's data path: profile words, selector, S = H(diag) + signed profile (32-bit add)
            const uint32_t px = profile_word((uint32_t)tcode[0], k, 0), py = profile_word((uint32_t)tcode[1], k, 1);
            const uint32_t ca = (uint32_t)qcode[0], cb = (uint32_t)qcode[1];
            const uint32_t qsel = ca | ((8u | ca) << 4) | ((4u + cb) << 8) | ((12u + cb) << 12);
            uint32_t subw = prmt(px, py, qsel);
            const uint32_t nmask = (qn[0] ? 0u : 0x0000FFFFu) | (qn[1] ? 0u : 0xFFFF0000u);
            subw = bitsel(nmask, subw, k.sub_n);
            const uint32_t Hd = (uint32_t)(diag[0] + kBias) | ((uint32_t)(diag[1] + kBias) << 16);
            const uint32_t s = Hd + subw;
            const uint32_t Fw = (uint32_t)(F[0] + kBias) | ((uint32_t)(F[1] + kBias) << 16);
            const uint32_t ew = (uint32_t)(e[0] + kBias) | ((uint32_t)(e[1] + kBias) << 16);
            uint32_t h, fn, en, nib, key;
            fast_cell(k, s, Fw, ew, key_colconst<5>(7), k.k32, h, fn, en, nib, key);
            for (int p = 0; p < 2; ++p) {
                const Ref r = ref_cell(sc, diag[p], F[p], e[p], sub[p]);
                const int gh = half_s(h, p) - kBias, gf = half_s(fn, p) - kBias, ge = half_s(en, p) - kBias;
                const int gn = (int)((nib >> (16 * p)) & 0xFFFFu);
                const int gk = half_s(key, p);
                ++checked;
                if (gh != r.h || gf != r.fn || ge != r.en || gn != r.nib || (r.h < 1024 && gk != ((r.h << 5) | (31 - 7)))) {
                    if (++bad <= 10)
                        fprintf(stderr, "mismatch sc=(%d,%d,%d,%d) half %d: diag %d F %d e %d sub %d -> h %d/%d fn %d/%d en %d/%d nib %x/%x key %x\n",
                                sc.match, sc.mismatch, sc.gap_oe, sc.gap_ext, p, diag[p], F[p], e[p], sub[p], gh, r.h, gf, r.fn, ge, r.en, gn, r.nib, gk);
                }
            }
        }
        // the gather: every combination of four nibbles per half
        for (uint32_t v = 0; v < 65536; ++v) {
            const uint32_t w = v ^ 0x5A5Au;   // pair B's nibbles
            uint32_t n[4];
            for (int c = 0; c < 4; ++c) n[c] = ((v >> (4 * c)) & 0xFu) | (((w >> (4 * c)) & 0xFu) << 16);
            const uint32_t word = dir_word(dir_pair(k, n[0], n[1]), dir_pair(k, n[2], n[3]));
            const uint32_t w3 = dir_word(dir_pair(k, n[0], n[1]), n[2]);
            ++checked;
            if (word != (v | (w << 16)) || w3 != ((v & 0x0FFFu) | ((w & 0x0FFFu) << 16))) {
                if (++bad <= 10) fprintf(stderr, "gather mismatch %04x/%04x -> %08x %08x\n", v, w, word, w3);
            }
        }
    }
    printf("{\"checked\": %lld, \"bad\": %lld}\n", checked, bad);
    return bad ? 1 : 0;
}
