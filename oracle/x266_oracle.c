/*
 * oracle/x266_oracle.c -- CPU restatement of the x266 block-parallel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by, or executed from the
 * product library (x266_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / reported baseline.
 *
 * Every function restates (in its own words, not copied) the algorithm of a reference function and
 * cites it.  Reference = chenm001/x266 @ 379268c, paths relative to the reference root.
 *
 * Pinning status (see tests/test_oracle.py):
 *   - DCT32 / SATD8x8: PINNED against the unmodified reference built in place (oracle/_ref, from
 *     src_tb/dct32.c + src_tb/satd.c) on random, extreme and wrap-around inputs, and against the
 *     known-answer hashes of SURVEY.md section 8(c) (tests/golden/kat.json).
 *   - DCT 4/8/16: PINNED *through the reference code* by the palindromic-extension identity
 *     (a size-N row repeated as [x, rev x, ...] to 32 samples through the reference
 *     partialButterfly32 with shift + (5-log2N) equals the size-N transform on rows k*32/N).
 *   - Intra32: NO EXECUTABLE REFERENCE exists (no C model; src/mkIntra32-wip.bsv does not compile), so
 *     no prediction of the reference can be reproduced and the header keeps saying "parity unpinned"
 *     for the pixels.  Pinned as far as the reference goes: every table (mapTbl, facTbl, mapShift,
 *     all rows, by the modes their comments name), every reference-line / inverse-angle projection
 *     list of getRefPixels (all 18 cases), the interpolator weights and the DC sum are parsed from
 *     the BSV into tests/golden/intra_bsv.json and this restatement reproduces each; the WIP file's
 *     defects are listed in the tests as explicit expected differences (two swapped mapTbl entries
 *     of mode 16, the undefined xL[0] in the corner slot and xT[32] for xL[32] in cases 9-15, the
 *     live >>6, the truncating and doubled DC shift).  A second, table-free restatement
 *     (orc_intra32_direct) backs the first on all 35 modes.
 *   - SATD full search / SAD: the search loop is ours (reference has no search loop); the cost of
 *     each candidate is the pinned satd8x8.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------------
 * A1. Transform matrix.  src_tb/dct32.c:30-64 (g_t32) and src/mkDct32.bsv:39-73 (left half).
 * Restated from its structure instead of as a literal table: entry (k,n) is the integerised cosine
 * c[a] with a = k*(2n+1) taken modulo 128 (angle a*pi/64) and reflected into [0,32]; k = 0 is the
 * flat row of 64.  The 33 magnitudes are the distinct values of the HEVC/VVC DCT-II matrix.
 * ---------------------------------------------------------------------------------------------- */
static const short orc_cos64[33] = {
    /* a = 0..32 ; a=0 is never used for k>0 because k*(2n+1) is never 0 mod 128 */
    91, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
    61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0
};

static int orc_g(int k, int n)
{
    int a;
    if (k == 0) return 64;
    a = (k * (2 * n + 1)) & 127;
    if (a <= 32) return orc_cos64[a];
    if (a <= 64) return -orc_cos64[64 - a];
    if (a <= 96) return -orc_cos64[a - 64];
    return orc_cos64[128 - a];
}

void orc_build_g32(short g[32 * 32])
{
    int k, n;
    for (k = 0; k < 32; k++)
        for (n = 0; n < 32; n++)
            g[k * 32 + n] = (short)orc_g(k, n);
}

/* ------------------------------------------------------------------------------------------------
 * A2. One 1-D pass with transposed store.  src_tb/dct32.c:66-170 (partialButterfly32).
 *   dst[k*line + j] = (int16)((sum_n G_N[k][n] * src[j*N + n] + (1 << (shift-1))) >> shift)
 * with G_N[k][n] = g_t32[k*32/N][n] (the sub-butterflies at dct32.c:109-143 index g_t32 that way),
 * 32-bit accumulation, arithmetic shift, truncating (mod 2^16) cast -- dct32.c:128-151.
 * Restated as the even/odd recursion the reference unrolls by hand (E/O :78-82, EE/EO :109-113,
 * EEE/EEO :116-120, EEEE/EEEO :123-126): even outputs are the half-size transform of the folded
 * sums, odd outputs are a dense half-length dot product with the folded differences.
 * ---------------------------------------------------------------------------------------------- */
static void orc_dct1d_rec(const int* x, int n, int step, int* out /* out[k*step] for k<n */)
{
    int e[16], o[16];
    int half = n >> 1, k, m;
    if (n == 1) { out[0] = 64 * x[0]; return; }           /* g_t32[0][0] */
    for (m = 0; m < half; m++) { e[m] = x[m] + x[n - 1 - m]; o[m] = x[m] - x[n - 1 - m]; }
    orc_dct1d_rec(e, half, step * 2, out);                 /* even k */
    for (k = 1; k < n; k += 2)
    {
        int acc = 0;
        for (m = 0; m < half; m++) acc += orc_g(k * (32 / n), m) * o[m];
        out[k * step] = acc;
    }
}

void orc_partialButterfly(const int16_t* src, int16_t* dst, int shift, int line, int log2n)
{
    const int n = 1 << log2n;
    const int add = 1 << (shift - 1);
    int j, k, x[32], y[32];
    for (j = 0; j < line; j++)
    {
        for (k = 0; k < n; k++) x[k] = src[j * n + k];
        orc_dct1d_rec(x, n, 1, y);
        for (k = 0; k < n; k++) dst[k * line + j] = (int16_t)((y[k] + add) >> shift);
    }
}

/* Independent second restatement: the dense matrix product (no butterfly).  Used to cross-check the
 * recursion above and as the spec the tensor-core kernel implements. */
void orc_partialDense(const int16_t* src, int16_t* dst, int shift, int line, int log2n)
{
    const int n = 1 << log2n;
    const int add = 1 << (shift - 1);
    int j, k, m;
    for (j = 0; j < line; j++)
        for (k = 0; k < n; k++)
        {
            int acc = 0;
            for (m = 0; m < n; m++) acc += orc_g(k * (32 / n), m) * src[j * n + m];
            dst[k * line + j] = (int16_t)((acc + add) >> shift);
        }
}

/* A3. 2-D composition.  src_tb/dct32.c:197-198: two passes, the two transposed stores cancel. */
void orc_dct2d(const int16_t* src, int16_t* dst, int log2n, int shift1, int shift2)
{
    int16_t coef[32 * 32];
    const int n = 1 << log2n;
    orc_partialButterfly(src, coef, shift1, n, log2n);
    orc_partialButterfly(coef, dst, shift2, n, log2n);
}

/* "Next" row N3: inverse transform (PARITY UNPINNED: no inverse exists in the reference).  Defined as the HEVC/VVC
 * decoder and the HM model partialButterflyInverse32 do it: for every column j of the stored block,
 *   dst[j*N + k] = clip16((sum_u G_N[u][k] * src[u*line + j] + (1 << (shift-1))) >> shift)
 * (transposed store, saturating), applied twice. */
static int16_t orc_clip16(int v) { return (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

void orc_partialInverse(const int16_t* src, int16_t* dst, int shift, int line, int log2n)
{
    const int n = 1 << log2n, add = 1 << (shift - 1);
    int j, k, u;
    for (j = 0; j < line; j++)
        for (k = 0; k < n; k++)
        {
            int acc = 0;
            for (u = 0; u < n; u++) acc += orc_g(u * (32 / n), k) * src[u * line + j];
            dst[j * n + k] = orc_clip16((acc + add) >> shift);
        }
}

void orc_idct2d(const int16_t* src, int16_t* dst, int log2n, int shift1, int shift2)
{
    int16_t tmp[32 * 32];
    const int n = 1 << log2n;
    orc_partialInverse(src, tmp, shift1, n, log2n);
    orc_partialInverse(tmp, dst, shift2, n, log2n);
}

/* ------------------------------------------------------------------------------------------------
 * A4. 8x8 Hadamard SATD.  src_tb/satd.c:31-118 (satd8x8).
 * Rows then columns, each a 3-stage Hadamard with partner distances 4, 2, 1 (satd.c:41-66 and
 * :73-100); every stage result is stored to int16 (satd.c:35) so it wraps mod 2^16; the cost is
 * (sum |coef| + 2) >> 2 in 32 bits (satd.c:105-113).
 * ---------------------------------------------------------------------------------------------- */
static void orc_had8(int16_t* v, int stride)
{
    int dist, i;
    for (dist = 4; dist >= 1; dist >>= 1)
    {
        int16_t t[8];
        int pos = 0;
        /* outputs are ordered sums-then-differences inside each group of 2*dist, like the
         * reference; any consistent order gives the same multiset of final coefficients */
        for (i = 0; i < 8; i++)
        {
            int grp = i / (2 * dist), idx = i % (2 * dist);
            int a = grp * 2 * dist + (idx % dist);
            int b = a + dist;
            int s = (idx < dist) ? (v[a * stride] + v[b * stride]) : (v[a * stride] - v[b * stride]);
            t[pos++] = (int16_t)s;                          /* int16 store == wrap */
        }
        for (i = 0; i < 8; i++) v[i * stride] = t[i];
    }
}

int orc_satd8x8(const int16_t diff[64])
{
    int16_t m[64];
    int32_t sad = 0;
    int i;
    memcpy(m, diff, sizeof(m));
    for (i = 0; i < 8; i++) orc_had8(m + 8 * i, 1);         /* horizontal, satd.c:39-70 */
    for (i = 0; i < 8; i++) orc_had8(m + i, 8);             /* vertical,   satd.c:73-103 */
    for (i = 0; i < 64; i++) sad += abs((int)m[i]);
    return (sad + 2) >> 2;
}

/* ------------------------------------------------------------------------------------------------
 * SATD full search (the window convention is ours, the reference has no
 * search loop).  cur: w x h u8 plane, stride w.  refPad: reference plane edge-replicated by `range`
 * pixels on every side, stride strd; pixel (x,y) of the unpadded plane is refPad[(y+range)*strd +
 * x+range].  For the 8x8 block whose top-left is (bx,by) and mv=(mvx,mvy) in [-range,range]^2:
 *   diff[y][x] = cur[by+y][bx+x] - ref[by+mvy+y][bx+mvx+x]   (int16)
 *   cost[(mvy+range)*(2*range+1) + (mvx+range)] = satd8x8(diff)
 * ---------------------------------------------------------------------------------------------- */
void orc_satd_search_block(const uint8_t* cur, int w, const uint8_t* refPad, intptr_t strd,
                           int bx, int by, int range, uint32_t* cost)
{
    const int side = 2 * range + 1;
    int mvx, mvy, x, y;
    for (mvy = -range; mvy <= range; mvy++)
        for (mvx = -range; mvx <= range; mvx++)
        {
            int16_t d[64];
            for (y = 0; y < 8; y++)
                for (x = 0; x < 8; x++)
                    d[y * 8 + x] = (int16_t)((int)cur[(size_t)(by + y) * w + bx + x] -
                                             (int)refPad[(intptr_t)(by + mvy + y + range) * strd + bx + mvx + x + range]);
            cost[(mvy + range) * side + (mvx + range)] = (uint32_t)orc_satd8x8(d);
        }
}

/* argmin rule: lowest cost; ties -> smallest mvx^2+mvy^2; then raster order (mvy, then mvx). */
void orc_satd_argmin(const uint32_t* cost, int range, uint32_t* bestCost, int* bestMvx, int* bestMvy)
{
    const int side = 2 * range + 1;
    uint32_t bc = 0xFFFFFFFFu; int bd = 0x7FFFFFFF, bxm = 0, bym = 0, ix, iy;
    for (iy = 0; iy < side; iy++)
        for (ix = 0; ix < side; ix++)
        {
            uint32_t c = cost[iy * side + ix];
            int mx = ix - range, my = iy - range, d = mx * mx + my * my;
            if (c < bc || (c == bc && d < bd)) { bc = c; bd = d; bxm = mx; bym = my; }
        }
    *bestCost = bc; *bestMvx = bxm; *bestMvy = bym;
}

/* whole frame; cost may be NULL (argmin only); best = [nBlocks][3] = {cost, mvx, mvy} or NULL */
void orc_satd_search_frame(const uint8_t* cur, int w, int h, const uint8_t* refPad, intptr_t strd,
                           int range, size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    const int bw = w / 8;
    const size_t side = (size_t)(2 * range + 1);
    uint32_t* tmp = (uint32_t*)malloc(side * side * sizeof(uint32_t));
    size_t b;
    (void)h;
    for (b = blk0; b < blk1; b++)
    {
        int bx = (int)(b % bw) * 8, by = (int)(b / bw) * 8;
        uint32_t* c = cost ? cost + (b - blk0) * side * side : tmp;
        orc_satd_search_block(cur, w, refPad, strd, bx, by, range, c);
        if (best)
        {
            uint32_t bc; int mx, my;
            orc_satd_argmin(c, range, &bc, &mx, &my);
            best[(b - blk0) * 3 + 0] = (int32_t)bc;
            best[(b - blk0) * 3 + 1] = mx;
            best[(b - blk0) * 3 + 2] = my;
        }
    }
    free(tmp);
}

/* Plain SAD of a w x h region (riscv/programs/benchmarks/sad/sad.c:27-38). */
uint32_t orc_sad(const uint8_t* a, intptr_t sa, const uint8_t* b, intptr_t sb, int w, int h)
{
    uint32_t s = 0; int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            s += (uint32_t)abs((int)a[y * sa + x] - (int)b[y * sb + x]);
    return s;
}

/* SAD full search: same window / argmin conventions as the SATD search, cost = orc_sad of the 8x8 block pair. */
void orc_sad_search_frame(const uint8_t* cur, int w, int h, const uint8_t* refPad, intptr_t strd,
                          int range, size_t blk0, size_t blk1, uint32_t* cost, int32_t* best)
{
    const int bw = w / 8, side = 2 * range + 1;
    uint32_t* tmp = (uint32_t*)malloc((size_t)side * side * sizeof(uint32_t));
    size_t b;
    (void)h;
    for (b = blk0; b < blk1; b++)
    {
        const int bx = (int)(b % bw) * 8, by = (int)(b / bw) * 8;
        uint32_t* c = cost ? cost + (b - blk0) * side * side : tmp;
        int mvx, mvy;
        for (mvy = -range; mvy <= range; mvy++)
            for (mvx = -range; mvx <= range; mvx++)
                c[(mvy + range) * side + mvx + range] =
                    orc_sad(cur + (size_t)by * w + bx, w, refPad + (intptr_t)(by + mvy + range) * strd + bx + mvx + range, strd, 8, 8);
        if (best)
        {
            uint32_t bc; int mx, my;
            orc_satd_argmin(c, range, &bc, &mx, &my);
            best[(b - blk0) * 3 + 0] = (int32_t)bc; best[(b - blk0) * 3 + 1] = mx; best[(b - blk0) * 3 + 2] = my;
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * A6. 32x32 intra prediction, 35 modes (0 planar, 1 DC, 2..34 angular; 10 = horizontal, 26 =
 * vertical).  PARITY UNPINNED (see header).  Spec source: src/mkIntra32-wip.bsv
 *   - refs: left[64] (left[i] = pixel (-1, i)), top[65] (top[0] = corner (-1,-1), top[1+i] =
 *     pixel (i,-1))                                           :34-37
 *   - per-row index / fraction tables = ((k+1)*angle)>>5 and &31 :75-112
 *   - negative-angle modes extend the main reference by projecting the side reference with
 *     (k*invAngle+128)>>8                                     :151-220
 *   - 2-tap interpolation ((32-f)*a + f*b + 16) >> 5 (the live rule has >>6 at :364, a WIP
 *     defect; the disabled block uses roundN(.,5) at :503)    :352-368
 *   - DC = (sum of 32 left + 32 top + 32) >> 6 (the RTL truncates without the +32 and shifts twice,
 *     :388-392 then :318; we use the standard rounding and say so)
 * No reference smoothing and no boundary filters (the RTL has none; for 32x32 the standard disables
 * the DC/angular edge filters anyway).
 * ---------------------------------------------------------------------------------------------- */
static const int orc_intra_angle[17]    = { 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32 };
static const int orc_intra_invAngle[8]  = { 4096, 1638, 910, 630, 482, 390, 315, 256 }; /* for angle -2..-32 */

int orc_intra_mode_angle(int mode)          /* mode 2..34 */
{
    return (mode >= 18) ? -orc_intra_angle[mode - 18] : orc_intra_angle[mode - 2];
    /* modes 2..18 walk +32..-32 (horizontal family, mirrored), 18..34 walk -32..+32 */
}

/* per-distance index / fraction of an angular mode: k = 0..31 is the distance from the main reference (row for the vertical
 * family, column for the horizontal one) -- mapTbl / facTbl of the BSV (:75-112) */
void orc_intra_idx_frac(int mode, int k, int* idx, int* frac)
{
    const int t = (k + 1) * orc_intra_mode_angle(mode);
    *idx = t >> 5;
    *frac = t & 31;
}

/* The working reference line of an angular mode, ref[-32..64] stored at out[0..96] (out[32] = ref[0] = the corner); entries the mode
 * never reads are -1.  Main part = top (vertical family) or corner + left (horizontal family); negative-angle modes extend it
 * to the left by projecting the side reference with (k*invAngle+128)>>8 -- getRefPixels of the BSV (:135-328). */
void orc_intra_ref_line(const uint8_t left[64], const uint8_t top[65], int mode, int out[97])
{
    const int isVer = mode >= 18;
    const int ang = orc_intra_mode_angle(mode);
    int *ref = out + 32, k;
    for (k = 0; k < 97; k++) out[k] = -1;
    for (k = 0; k <= 64; k++) ref[k] = isVer ? top[k] : (k == 0 ? top[0] : left[k - 1]);
    if (ang < 0)
    {
        int a, inv = 0;
        for (k = 33; k <= 64; k++) ref[k] = -1;          /* a negative angle never looks past ref[32] */
        for (a = 0; a < 8; a++) if (orc_intra_angle[9 + a] == ang) inv = orc_intra_invAngle[a];
        for (k = -1; k > ((32 * ang) >> 5); k--)          /* lowest index read: x + idx + 1 with x = 0, idx = (32*ang)>>5 */
        {
            int s = ((-k) * inv + 128) >> 8;             /* side-reference index 1..32 */
            ref[k] = isVer ? left[s - 1] : top[s];
        }
    }
}

void orc_intra32(const uint8_t left[64], const uint8_t top[65], int mode, uint8_t pred[32 * 32])
{
    int x, y;
    if (mode == 0)                           /* planar */
    {
        const int tr = top[1 + 32], bl = left[32];
        for (y = 0; y < 32; y++)
            for (x = 0; x < 32; x++)
                pred[y * 32 + x] = (uint8_t)(((31 - x) * left[y] + (x + 1) * tr +
                                              (31 - y) * top[1 + x] + (y + 1) * bl + 32) >> 6);
        return;
    }
    if (mode == 1)                           /* DC */
    {
        int s = 32;
        for (x = 0; x < 32; x++) s += left[x] + top[1 + x];
        memset(pred, s >> 6, 32 * 32);
        return;
    }
    {
        const int isVer = mode >= 18;
        int buf[97], *ref = buf + 32;
        orc_intra_ref_line(left, top, mode, buf);
        for (y = 0; y < 32; y++)                          /* y = distance from the main reference */
        {
            int idx, f;
            orc_intra_idx_frac(mode, y, &idx, &f);
            for (x = 0; x < 32; x++)
            {
                int v = f ? (((32 - f) * ref[x + idx + 1] + f * ref[x + idx + 2] + 16) >> 5)
                          : ref[x + idx + 1];
                if (isVer) pred[y * 32 + x] = (uint8_t)v; else pred[x * 32 + y] = (uint8_t)v;
            }
        }
    }
}

/* Second, independent restatement of the same predictor (as orc_partialDense backs the butterfly): one pixel at a time, straight
 * from the sample-position form of the 35-mode definition, no working line and no tables -- p(x, y) below is the neighbouring
 * sample at picture position (x, y) relative to the block, x = -1 or y = -1. */
static int orc_nb(const uint8_t left[64], const uint8_t top[65], int x, int y)
{
    return y < 0 ? top[x + 1] : left[y];               /* (x, -1) for x = -1..63 | (-1, y) for y = 0..63 */
}

static int orc_intra_sample(const uint8_t left[64], const uint8_t top[65], int isVer, int angle, int i)
{
    /* reference sample number i along the main direction: i >= 0 on the main side (i = 0 is the corner), i < 0 projected from
     * the side with the inverse angle round(8192 / -angle) */
    if (i >= 0) return isVer ? orc_nb(left, top, i - 1, -1) : (i == 0 ? orc_nb(left, top, -1, -1) : orc_nb(left, top, -1, i - 1));
    {
        const int inv = (8192 * 2 + (-angle)) / (2 * (-angle));       /* rounded 8192/|angle|: 4096 1638 910 630 482 390 315 256 */
        const int s = ((-i) * inv + 128) >> 8;
        return isVer ? orc_nb(left, top, -1, s - 1) : orc_nb(left, top, s - 1, -1);
    }
}

void orc_intra32_direct(const uint8_t left[64], const uint8_t top[65], int mode, uint8_t pred[32 * 32])
{
    static const signed char angle_of_mode[35] = { 0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26,
                                                   -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
    int px, py;
    for (py = 0; py < 32; py++)
        for (px = 0; px < 32; px++)
        {
            int v;
            if (mode == 0)
                v = ((31 - px) * orc_nb(left, top, -1, py) + (px + 1) * orc_nb(left, top, 32, -1) +
                     (31 - py) * orc_nb(left, top, px, -1) + (py + 1) * orc_nb(left, top, -1, 32) + 32) >> 6;
            else if (mode == 1)
            {
                int i, s = 0;
                for (i = 0; i < 32; i++) s += orc_nb(left, top, i, -1) + orc_nb(left, top, -1, i);
                v = (s + 32) >> 6;
            }
            else
            {
                const int isVer = mode >= 18, angle = angle_of_mode[mode];
                const int dist = isVer ? py : px, along = isVer ? px : py;     /* distance from / position along the main reference */
                const int disp = (dist + 1) * angle;                           /* displacement in 1/32 sample */
                const int whole = disp >= 0 ? disp / 32 : -((-disp + 31) / 32); /* floor without relying on >> of a negative */
                const int part = disp - 32 * whole;
                const int a = orc_intra_sample(left, top, isVer, angle, along + whole + 1);
                v = a;
                if (part)
                    v = ((32 - part) * a + part * orc_intra_sample(left, top, isVer, angle, along + whole + 2) + 16) / 32;
            }
            pred[py * 32 + px] = (uint8_t)v;
        }
}

/* Fused intra mode decision ("next" row N1): cost[m] = sum over the 16 8x8 sub-blocks of satd8x8(cur - pred_m). */
void orc_intra32_decide(const uint8_t cur[1024], const uint8_t left[64], const uint8_t top[65], uint32_t cost[35], int32_t* bestMode)
{
    int m, sb, x, y, bm = 0;
    for (m = 0; m < 35; m++)
    {
        uint8_t pred[1024];
        uint32_t c = 0;
        orc_intra32(left, top, m, pred);
        for (sb = 0; sb < 16; sb++)
        {
            int16_t d[64];
            for (y = 0; y < 8; y++)
                for (x = 0; x < 8; x++)
                {
                    const int o = ((sb >> 2) * 8 + y) * 32 + (sb & 3) * 8 + x;
                    d[y * 8 + x] = (int16_t)((int)cur[o] - (int)pred[o]);
                }
            c += (uint32_t)orc_satd8x8(d);
        }
        cost[m] = c;
        if (c < cost[bm]) bm = m;
    }
    *bestMode = bm;
}

/* ------------------------------------------------------------------------------------------------
 * Tiled frame format ("next" row N2).  src/x266.cpp:56-63 (ref_block_t: 256 B luma 16x16, 128 B chroma as 8
 * rows of 8 (U,V) pairs, 128 B info) and src/x266.cpp:415-492 (xConvInputFmt / xConvOutput420): tiles in
 * raster order, width/16 per row.  Restated per pixel instead of per memcpy row.
 * ---------------------------------------------------------------------------------------------- */
void orc_conv_input_fmt(uint8_t* tiles, const uint8_t* Y, const uint8_t* U, const uint8_t* V, intptr_t strdY, int w, int h)
{
    const intptr_t strdC = strdY >> 1;
    const int tpr = w / 16;
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            tiles[((size_t)(y / 16) * tpr + x / 16) * 512 + (y % 16) * 16 + (x % 16)] = Y[y * strdY + x];
    for (y = 0; y < h / 2; y++)
        for (x = 0; x < w / 2; x++)
        {
            uint8_t* t = tiles + ((size_t)(y / 8) * tpr + x / 8) * 512 + 256 + (y % 8) * 16 + (x % 8) * 2;
            t[0] = U[y * strdC + x];
            t[1] = V[y * strdC + x];
        }
}

void orc_conv_output420(const uint8_t* tiles, uint8_t* Y, intptr_t strdY, uint8_t* U, uint8_t* V, intptr_t strdC, int w, int h)
{
    const int tpr = w / 16;
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            Y[y * strdY + x] = tiles[((size_t)(y / 16) * tpr + x / 16) * 512 + (y % 16) * 16 + (x % 16)];
    for (y = 0; y < h / 2; y++)
        for (x = 0; x < w / 2; x++)
        {
            const uint8_t* t = tiles + ((size_t)(y / 8) * tpr + x / 8) * 512 + 256 + (y % 8) * 16 + (x % 8) * 2;
            U[y * strdC + x] = t[0];
            V[y * strdC + x] = t[1];
        }
}

/* luma residual of 32x32 block b (raster) between two tiled frames, then the pinned 2-D transform */
void orc_frame_resi_dct32(const uint8_t* cur, const uint8_t* pred, int w, int h, int16_t* coef, int s1, int s2)
{
    const int tpr = w / 16, bpr = w / 32;
    size_t b, nb = (size_t)bpr * (h / 32);
    for (b = 0; b < nb; b++)
    {
        int16_t r[1024];
        int x, y;
        const int bx = (int)(b % bpr) * 32, by = (int)(b / bpr) * 32;
        for (y = 0; y < 32; y++)
            for (x = 0; x < 32; x++)
            {
                const size_t o = ((size_t)((by + y) / 16) * tpr + (bx + x) / 16) * 512 + ((by + y) % 16) * 16 + ((bx + x) % 16);
                r[y * 32 + x] = (int16_t)((int)cur[o] - (int)pred[o]);
            }
        orc_dct2d(r, coef + b * 1024, 5, s1, s2);
    }
}

/* ------------------------------------------------------------------------------------------------
 * KAT helpers: splitmix64 stream and FNV-1a-64.
 * ---------------------------------------------------------------------------------------------- */
uint64_t orc_fnv1a64(const void* buf, size_t n)
{
    const uint8_t* p = (const uint8_t*)buf;
    uint64_t h = 1469598103934665603ull;
    size_t i;
    for (i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static uint64_t orc_sm64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* kind 0: 9-bit (z&0xFF)-((z>>8)&0xFF); 1: 11-bit (z&0x3FF)-((z>>10)&0x3FF); 2: full int16 */
void orc_fill_residual(int16_t* dst, size_t n, uint64_t seed, int kind)
{
    uint64_t s = seed; size_t i;
    for (i = 0; i < n; i++)
    {
        uint64_t z = orc_sm64(&s);
        dst[i] = kind == 0 ? (int16_t)((int)(z & 0xFF) - (int)((z >> 8) & 0xFF))
               : kind == 1 ? (int16_t)((int)(z & 0x3FF) - (int)((z >> 10) & 0x3FF))
                           : (int16_t)(z & 0xFFFF);
    }
}

/* ------------------------------------------------------------------------------------------------
 * N3 quantiser stub + N1/N3 closed block loop.  NOT IN THE REFERENCE: the arithmetic is that of the HEVC/VVC test
 * models' flat intra quantiser for 8-bit video and a 32x32 transform (transformShift = 15 - 8 - 5 = 2):
 *   level = sign(c) * min(32767, (|c| * qScale[qp%6] + (171 << (qBits-9))) >> qBits),  qBits = 14 + qp/6 + 2
 *   c'    = clip16((level * (iqScale[qp%6] << (qp/6)) + 8) >> 4)
 * and the loop is  decide (orc_intra32_decide) -> best-mode prediction (orc_intra32) -> residual -> orc_dct2d(4, 11) -> quantise
 * -> de-quantise -> orc_idct2d(7, 12) -> recon = clip8(pred + residual').
 * ---------------------------------------------------------------------------------------------- */
static const int orc_qScale[6]  = { 26214, 23302, 20560, 18396, 16384, 14564 };
static const int orc_iqScale[6] = { 40, 45, 51, 57, 64, 72 };

void orc_quant(const int16_t* coef, int16_t* level, size_t n, int qp)
{
    const int qBits = 16 + qp / 6;
    const int64_t add = (int64_t)171 << (qBits - 9);
    size_t i;
    for (i = 0; i < n; i++)
    {
        const int c = coef[i];
        int64_t a = (((int64_t)(c < 0 ? -c : c) * orc_qScale[qp % 6] + add) >> qBits);
        if (a > 32767) a = 32767;
        level[i] = (int16_t)(c < 0 ? -a : a);
    }
}

void orc_dequant(const int16_t* level, int16_t* coef, size_t n, int qp)
{
    const int64_t scale = (int64_t)orc_iqScale[qp % 6] << (qp / 6);
    size_t i;
    for (i = 0; i < n; i++)
    {
        int64_t v = (level[i] * scale + 8) >> 4;
        coef[i] = (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v);
    }
}

/* the `Recon` channel (mkIntra32-wip.bsv:39-48): mode given */
void orc_intra32_recon(const uint8_t cur[1024], const uint8_t left[64], const uint8_t top[65], int mode, int qp,
                       int16_t level[1024], uint8_t recon[1024])
{
    uint8_t pred[1024];
    int16_t resi[1024], coef[1024], dq[1024], back[1024];
    int i;
    orc_intra32(left, top, mode, pred);
    for (i = 0; i < 1024; i++) resi[i] = (int16_t)(cur[i] - pred[i]);
    orc_dct2d(resi, coef, 5, 4, 11);
    orc_quant(coef, level, 1024, qp);
    orc_dequant(level, dq, 1024, qp);
    orc_idct2d(dq, back, 5, 7, 12);
    for (i = 0; i < 1024; i++)
    {
        const int v = pred[i] + back[i];
        recon[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

void orc_intra32_encode(const uint8_t cur[1024], const uint8_t left[64], const uint8_t top[65], int qp,
                        int16_t level[1024], uint8_t recon[1024], uint32_t cost[35], int32_t* bestMode)
{
    orc_intra32_decide(cur, left, top, cost, bestMode);
    orc_intra32_recon(cur, left, top, *bestMode, qp, level, recon);
}

/* ------------------------------------------------------------------------------------------------
 * Batch drivers with a pthread fan-out over contiguous ranges (CPU baseline "port" when the real
 * reference library oracle/_ref is not available).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { const int16_t* src; void* dst; size_t lo, hi; int a, b, c; int kind; } orc_job_t;

static void* orc_worker(void* p)
{
    orc_job_t* j = (orc_job_t*)p; size_t i;
    if (j->kind == 0)
    {
        const size_t bs = (size_t)1 << (2 * j->a);
        for (i = j->lo; i < j->hi; i++)
            orc_dct2d(j->src + i * bs, (int16_t*)j->dst + i * bs, j->a, j->b, j->c);
    }
    else
        for (i = j->lo; i < j->hi; i++) ((int32_t*)j->dst)[i] = orc_satd8x8(j->src + i * 64);
    return 0;
}

static int orc_fanout(orc_job_t proto, size_t n, int threads)
{
    pthread_t th[1024]; orc_job_t jobs[1024]; int t;
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    for (t = 0; t < threads; t++)
    {
        jobs[t] = proto;
        jobs[t].lo = n * (size_t)t / threads;
        jobs[t].hi = n * (size_t)(t + 1) / threads;
    }
    if (threads == 1) { orc_worker(&jobs[0]); return 0; }
    for (t = 0; t < threads; t++) if (pthread_create(&th[t], 0, orc_worker, &jobs[t])) return -1;
    for (t = 0; t < threads; t++) pthread_join(th[t], 0);
    return 0;
}

int orc_dct_batch(const int16_t* src, int16_t* dst, size_t nBlocks, int log2n, int s1, int s2, int threads)
{
    orc_job_t p; memset(&p, 0, sizeof(p));
    p.src = src; p.dst = dst; p.a = log2n; p.b = s1; p.c = s2; p.kind = 0;
    return orc_fanout(p, nBlocks, threads);
}

int orc_satd8x8_batch(const int16_t* diff, int32_t* out, size_t n, int threads)
{
    orc_job_t p; memset(&p, 0, sizeof(p));
    p.src = diff; p.dst = out; p.kind = 1;
    return orc_fanout(p, n, threads);
}
