// CPU model of the packed/biased transform used by the search kernel v3: compiles the product's own inline
// arithmetic (x266_b200/csrc/satd_packed.h) with g++ so tests can compare it with the oracle without a GPU.
#include "../../x266_b200/csrc/satd_packed.h"

using namespace x266::s3;

// transform of one 8x8 pixel block (row-major, pitch 8) into the 32 packed words; returns the largest half seen
static uint32_t transform(const uint8_t* p, uint32_t T[32])
{
    uint32_t px[2][8], col[8][4];
    for (int i = 0; i < 8; i++) {
        px[0][i] = p[8 * i] | (p[8 * i + 1] << 8) | (p[8 * i + 2] << 16) | ((uint32_t)p[8 * i + 3] << 24);
        px[1][i] = p[8 * i + 4] | (p[8 * i + 5] << 8) | (p[8 * i + 6] << 16) | ((uint32_t)p[8 * i + 7] << 24);
    }
    vertical4(px[0], &col[0]);
    vertical4(px[1], &col[4]);
    for (int j = 0; j < 4; j++)
        for (int c = 0; c < 8; c++) T[8 * j + c] = col[c][j];
    horizontal8(T);
    uint32_t mx = 0;
    for (int k = 0; k < 32; k++) {
        mx = (T[k] & 0xFFFF) > mx ? (T[k] & 0xFFFF) : mx;
        mx = (T[k] >> 16) > mx ? (T[k] >> 16) : mx;
    }
    return mx;
}

extern "C" int packed_satd(const uint8_t* cur, const uint8_t* ref, int n, int32_t* cost, uint32_t* maxHalf)
{
    uint32_t worst = 0;
    for (int b = 0; b < n; b++) {
        uint32_t Tc[32], Tr[32];
        const uint32_t m0 = transform(cur + 64 * b, Tc), m1 = transform(ref + 64 * b, Tr);
        worst = m0 > worst ? m0 : worst;
        worst = m1 > worst ? m1 : worst;
        cost[b] = (int32_t)cost_from_maxsum(maxsum(Tr, Tc), ref[64 * b], cur[64 * b]);
    }
    *maxHalf = worst;
    return 0;
}
