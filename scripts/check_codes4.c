// Exhaustive host check of the PRMT-table ASCII -> 2-bit step (codes4 in rust-pseudoaligner_b200/csrc/psa_kernels.cuh):
// the same arithmetic with __byte_perm emulated, compared with DnaString::from_dna_string's mapping for all 2^32 inputs.
// gcc -O2 -o /tmp/codes4 scripts/check_codes4.c && /tmp/codes4   (about two minutes; prints bad = 0)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
static uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    uint64_t v = ((uint64_t)y << 32) | x; uint32_t r = 0;
    for (int i = 0; i < 4; i++) { uint32_t n = (s >> (4 * i)) & 7; r |= (uint32_t)((v >> (8 * n)) & 0xFF) << (8 * i); }
    return r;
}
static uint32_t codes4(uint32_t w) {
    uint32_t t = w & 0x07070707u; t |= t >> 4;
    uint32_t sel = byte_perm(t, 0u, 0x4420u);
    uint32_t x = byte_perm(0x01000000u, 0x02000003u, sel);
    uint32_t want = byte_perm(0x43FF41FFu, 0x47FFFF54u, sel);
    uint32_t d = (w & 0xDFDFDFDFu) ^ want;
    uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    x &= ~((nz >> 7) | (nz >> 6));
    return (x * 0x40100401u) >> 24;
}
static uint32_t ref1(uint8_t c) { c &= 0xDF; return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }
int main() {
    uint64_t bad = 0;
    for (uint64_t it = 0; it < (1ull << 32); it += 1) {
        uint32_t w = (uint32_t)it;
        if ((it & 0xFFFFFF) == 0 && 0) printf("%llu\n", (unsigned long long)it);
        uint32_t want = (ref1(w & 0xFF) << 6) | (ref1((w >> 8) & 0xFF) << 4) | (ref1((w >> 16) & 0xFF) << 2) | ref1(w >> 24);
        if (codes4(w) != want) { if (bad < 5) printf("bad %08x: %x vs %x\n", w, codes4(w), want); bad++; }
    }
    printf("bad = %llu\n", (unsigned long long)bad);
    return bad != 0;
}
