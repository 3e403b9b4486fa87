// chess_core.cuh -- device-side rules of the gym-chess v2 env on bitboards (sm_100a).
//
// Re-design, not a translation: the reference walks an 8x8 isize mailbox with HashMap
// attack maps and simulates every candidate move (src/lib.rs:501-677); here a position is
// four 64-bit planes (three piece-code bit-planes + a colour plane, 32 B), slider attacks
// come from Hyperbola-Quintessence with __brevll on tabulated line masks, the enemy attack
// map is one u64, the legality filter is a check mask + pin rays computed once per position,
// and the legal set is one 64-bit target set per own piece (the ordered list is a decode).
// Small geometry tables are built at compile time (constexpr) into device global memory.
// The OBSERVABLE behaviour (move order, quirks Q1-Q23 of SURVEY.md section 9) is that of
// the reference; each routine cites the lines whose results it reproduces.
//
// Square index = row*8 + col, row 0 = rank 8 (lib.rs:41-50, 1235-1238), so a row-major
// scan of the board (lib.rs:510-511) is an ascending bit scan.
#pragma once
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#ifndef GCB_ROLL_NTH
#define GCB_ROLL_NTH 1
#endif

// The rules are plain integer code; they are marked __host__ __device__ so that tests/host_emul can
// compile the SAME source with g++ and check the logic against the oracle on a box without a GPU.
// The product never runs them on the host (gym_chess_b200 has no CPU path).
#if defined(__CUDACC__)
#define GCB_HD __host__ __device__ __forceinline__
#else
#define GCB_HD inline
#endif

GCB_HD u64 gcb_brev64(u64 x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
GCB_HD int gcb_lsb(u64 x) {  // index of the lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
#ifndef GCB_BFIND  // device gcb_msb: one FLO on the non-zero half (1) or the compiler's 63 - clz64 (0)
#define GCB_BFIND 1
#endif
#ifndef GCB_TAKE_BELOW  // piece loops clear the taken square with the mask they need for the slot rank anyway
#define GCB_TAKE_BELOW 1
#endif
#ifndef GCB_ZONE_FILTER  // opponent attack map: only the sliders that can reach the king zone are evaluated
#define GCB_ZONE_FILTER 1
#endif
#ifndef GCB_NTH_TWO_LEVEL
#define GCB_NTH_TWO_LEVEL 1
#endif
#ifndef GCB_NTH_BSEARCH  // nth_target: binary search over cumulative direction masks instead of a walk over the 8 directions
#define GCB_NTH_BSEARCH 1
#endif
GCB_HD int gcb_msb(u64 x) {  // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__) && GCB_BFIND
    // FLO returns the bit index itself: pick the half, one FLO, add 32 for the high half (63 - clz64 compiles to three more
    // integer instructions, and this sits in every piece loop)
    const u32 hi = (u32)(x >> 32), lo = (u32)x;
    u32 r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(hi ? hi : lo));
    return (int)(hi ? r + 32u : r);
#elif defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)x);
#else
    return 63 - __builtin_clzll(x);
#endif
}
// take one square out of a non-empty set (the highest: FLO finds it without a bit reversal)
GCB_HD int gcb_take(u64& s) {
    const int sq = gcb_msb(s);
    s ^= 1ULL << sq;
    return sq;
}
// ... and also hand back the set of squares BELOW the taken one: the piece loops need it for the slot rank
// (popcount of the own pieces below), and since the taken square is the highest of s it clears the square as well
GCB_HD int gcb_take(u64& s, u64& below) {
    const int sq = gcb_msb(s);
#if GCB_TAKE_BELOW
    below = ~(~0ULL << sq);
    s &= below;
#else
    below = (1ULL << sq) - 1;
    s ^= 1ULL << sq;
#endif
    return sq;
}
GCB_HD int gcb_popc(u64 x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// index of the r-th (0-based) set bit of x, r < popc(x): a 6-step binary search on popcounts -- the same instructions in
// every lane, where clearing r low bits one by one would run the warp for the largest r
GCB_HD int gcb_select64(u64 x, int r) {
    u32 w = (u32)x;
    int base = 0, c = gcb_popc((u64)w);
    if (r >= c) r -= c, w = (u32)(x >> 32), base = 32;
    c = gcb_popc((u64)(w & 0xFFFFu));
    if (r >= c) r -= c, w >>= 16, base += 16;
    c = gcb_popc((u64)(w & 0xFFu));
    if (r >= c) r -= c, w >>= 8, base += 8;
    c = gcb_popc((u64)(w & 0xFu));
    if (r >= c) r -= c, w >>= 4, base += 4;
    c = gcb_popc((u64)(w & 0x3u));
    if (r >= c) r -= c, w >>= 2, base += 2;
    return base + ((w & 1u) ? r : 1);
}
GCB_HD u32 gcb_umulhi(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * (u64)b) >> 32);
#endif
}

#define GCB_FILE_A 0x0101010101010101ULL
#define GCB_FILE_B 0x0202020202020202ULL
#define GCB_FILE_G 0x4040404040404040ULL
#define GCB_FILE_H 0x8080808080808080ULL
#define GCB_MAIN_DIAG 0x8040201008040201ULL
#define GCB_ANTI_DIAG 0x0102040810204080ULL

// piece codes = reference ids (lib.rs:11-17)
enum { PC_NONE = 0, PC_KING = 1, PC_QUEEN = 2, PC_ROOK = 3, PC_BISHOP = 4, PC_KNIGHT = 5, PC_PAWN = 6 };

#define ACT_CASTLE_KS_WHITE 4096
#define ACT_CASTLE_QS_WHITE 4097
#define ACT_CASTLE_KS_BLACK 4098
#define ACT_CASTLE_QS_BLACK 4099
#define ACT_RESIGN 4100

// rights bits (wk, wq, bk, bq) = (*_king_castle_is_possible, *_queen_castle_is_possible)
#define RT_WK 1u
#define RT_WQ 2u
#define RT_BK 4u
#define RT_BQ 8u

struct Board {
    u64 t0, t1, t2;  // bit-planes of the piece code (K=001 Q=010 R=011 B=100 N=101 P=110)
    u64 w;           // colour plane: 1 = white piece
};

GCB_HD u64 bb_occ(const Board& b) { return b.t0 | b.t1 | b.t2; }
GCB_HD u64 bb_kings(const Board& b) { return b.t0 & ~b.t1 & ~b.t2; }
GCB_HD u64 bb_queens(const Board& b) { return ~b.t0 & b.t1 & ~b.t2; }
GCB_HD u64 bb_rooks(const Board& b) { return b.t0 & b.t1 & ~b.t2; }
GCB_HD u64 bb_bishops(const Board& b) { return ~b.t0 & ~b.t1 & b.t2; }
GCB_HD u64 bb_knights(const Board& b) { return b.t0 & ~b.t1 & b.t2; }
GCB_HD u64 bb_pawns(const Board& b) { return ~b.t0 & b.t1 & b.t2; }

GCB_HD int piece_code(const Board& b, int sq) {
    return (int)((b.t0 >> sq) & 1) | ((int)((b.t1 >> sq) & 1) << 1) | ((int)((b.t2 >> sq) & 1) << 2);
}
// signed reference id at sq (0 empty, + white, - black)
GCB_HD int piece_id(const Board& b, int sq) {
    int c = piece_code(b, sq);
    return ((b.w >> sq) & 1) ? c : -c;
}
GCB_HD void clear_sq(Board& b, u64 bit) {
    b.t0 &= ~bit, b.t1 &= ~bit, b.t2 &= ~bit, b.w &= ~bit;
}
GCB_HD void put_sq(Board& b, int sq, int code, int white) {
    u64 bit = 1ULL << sq;
    clear_sq(b, bit);
    if (code & 1) b.t0 |= bit;
    if (code & 2) b.t1 |= bit;
    if (code & 4) b.t2 |= bit;
    if (white && code) b.w |= bit;
}

// ---------------------------------------------------------------------------------------------
// Geometry tables, built at COMPILE time (constexpr) and placed in device global memory: every lookup is one
// L1-resident LDG instead of 10-20 integer instructions -- the step kernel is bound by the integer pipe, the
// load pipe is idle.  `ord` also is the single statement of the reference's per-piece move ORDER:
//   class 0 slider : rook dirs (-1,0)(1,0)(0,-1)(0,1), bishop dirs (-1,-1)(-1,1)(1,-1)(1,1)   lib.rs:824-851
//   class 1 king   : (1,0)(-1,0)(0,1)(0,-1)(1,1)(1,-1)(-1,1)(-1,-1)                            lib.rs:797-806
//   class 2 knight : (-2,-1)(-2,1)(2,-1)(2,1)(-1,-2)(-1,2)(1,-2)(1,2)                          lib.rs:891-900
//   class 3/4 white/black pawn : one step, two steps, (row-p, col+1), (row-p, col-1)            lib.rs:935-959
// Entry k of a piece on sq = the squares its k-th direction can reach on an empty board (a ray for sliders, one
// square for the others); the ordered move list of the piece is, for k = 0..7, (targets & entry k) by increasing
// distance.  Entries 0,2,4,5 of every class point to LOWER square indices (or hold a single square).
// ---------------------------------------------------------------------------------------------
struct alignas(16) GeomTables {
    u64 line[64][4];    // file, rank, diagonal, anti-diagonal through sq, WITHOUT sq
    u64 ord[5][64][8];  // ordered direction masks per class
    u64 knight[64], king[64];
    u64 pawn[2][64][2];  // [black][sq]: {push squares (one step; two from the start row), capture squares}
    u64 between[64][64]; // squares strictly between two aligned squares (0 when not aligned)
    u64 cum[5][64][8];   // cum[c][sq][k] = ord[c][sq][0] | ... | ord[c][sq][k]: nth_target's binary search
    // zone[0/1][k]: the squares from which a rook / bishop mover could reach, on an empty board, the king square k, one of
    // its neighbours or (k = e1 / e8) one of the castle squares -- the only places the opponent attack map is ever looked at
    u64 zone[2][64];
    constexpr GeomTables() : line(), ord(), knight(), king(), pawn(), between(), cum(), zone() {
        const int sl[8][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
        const int kg[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
        const int kn[8][2] = {{-2, -1}, {-2, 1}, {2, -1}, {2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}};
        const int wp[4][2] = {{-1, 0}, {-2, 0}, {-1, 1}, {-1, -1}};
        const int bp[4][2] = {{1, 0}, {2, 0}, {1, 1}, {1, -1}};
        for (int sq = 0; sq < 64; sq++) {
            const int r = sq >> 3, c = sq & 7;
            for (int t = 0; t < 64; t++) {
                const int tr = t >> 3, tc = t & 7;
                if (t == sq) continue;
                if (tc == c) line[sq][0] |= 1ULL << t;
                if (tr == r) line[sq][1] |= 1ULL << t;
                if (tr - tc == r - c) line[sq][2] |= 1ULL << t;
                if (tr + tc == r + c) line[sq][3] |= 1ULL << t;
            }
            for (int k = 0; k < 8; k++) {
                u64 walked = 0;
                for (int i = 1; i < 8; i++) {
                    const int tr = r + i * sl[k][0], tc = c + i * sl[k][1];
                    if (tr >= 0 && tr < 8 && tc >= 0 && tc < 8) {
                        ord[0][sq][k] |= 1ULL << (tr * 8 + tc);
                        between[sq][tr * 8 + tc] = walked;
                        walked |= 1ULL << (tr * 8 + tc);
                    }
                }
                int tr = r + kg[k][0], tc = c + kg[k][1];
                if (tr >= 0 && tr < 8 && tc >= 0 && tc < 8) ord[1][sq][k] |= 1ULL << (tr * 8 + tc), king[sq] |= 1ULL << (tr * 8 + tc);
                tr = r + kn[k][0], tc = c + kn[k][1];
                if (tr >= 0 && tr < 8 && tc >= 0 && tc < 8) ord[2][sq][k] |= 1ULL << (tr * 8 + tc), knight[sq] |= 1ULL << (tr * 8 + tc);
                if (k < 4) {
                    tr = r + wp[k][0], tc = c + wp[k][1];
                    if (tr >= 0 && tr < 8 && tc >= 0 && tc < 8) {
                        ord[3][sq][k] |= 1ULL << (tr * 8 + tc);
                        if (k != 1 || r == 6) pawn[0][sq][k >= 2] |= 1ULL << (tr * 8 + tc);  // two steps from row 6 only
                    }
                    tr = r + bp[k][0], tc = c + bp[k][1];
                    if (tr >= 0 && tr < 8 && tc >= 0 && tc < 8) {
                        ord[4][sq][k] |= 1ULL << (tr * 8 + tc);
                        if (k != 1 || r == 1) pawn[1][sq][k >= 2] |= 1ULL << (tr * 8 + tc);  // two steps from row 1 only
                    }
                }
            }
            for (int c5 = 0; c5 < 5; c5++) {
                u64 acc = 0;
                for (int k = 0; k < 8; k++) acc |= ord[c5][sq][k], cum[c5][sq][k] = acc;
            }
        }
        for (int k = 0; k < 64; k++) {  // (after the loop above: needs line[] and king[] of every square)
            u64 z = king[k] | (1ULL << k);
            if (k == 60) z |= 0x7CULL << 56;  // c1..g1 (gen_castles)
            if (k == 4) z |= 0x7CULL;         // c8..g8
            for (int t = 0; t < 64; t++)
                if ((z >> t) & 1) zone[0][k] |= line[t][0] | line[t][1] | (1ULL << t), zone[1][k] |= line[t][2] | line[t][3] | (1ULL << t);
        }
    }
};
#if defined(__CUDACC__)
__device__ const GeomTables g_geom_dev = GeomTables();
#endif
static const GeomTables g_geom_host = GeomTables();
#if defined(__CUDA_ARCH__)
#define GCB_GEOM(field) __ldg(&g_geom_dev.field)
#else
#define GCB_GEOM(field) (g_geom_host.field)
#endif
// Where the SMALL hot tables (line masks, knight / king sets, pawn spans: 5 KB) are read from.  GeomGlobal: the
// L1-resident device-memory copy.  GeomShared (device only): a copy the multi-step kernel keeps in shared memory for the
// whole launch -- LDS with 32-bit addresses instead of LDG with 64-bit address arithmetic on the integer pipe.
struct GeomGlobal {
    GCB_HD u64 line(int sq, int k) const { return GCB_GEOM(line[sq][k]); }
    GCB_HD u64 knight(int sq) const { return GCB_GEOM(knight[sq]); }
    GCB_HD u64 king(int sq) const { return GCB_GEOM(king[sq]); }
    GCB_HD u64 pawn(int black, int sq, int k) const { return GCB_GEOM(pawn[black][sq][k]); }
};
#define GCB_SGEOM_WORDS (64 * 4 + 64 + 64 + 2 * 64 * 2)
#if defined(__CUDACC__)
struct GeomShared {
    const u64* p;  // line[64][4] | knight[64] | king[64] | pawn[2][64][2]
    __device__ __forceinline__ u64 line(int sq, int k) const { return p[sq * 4 + k]; }
    __device__ __forceinline__ u64 knight(int sq) const { return p[256 + sq]; }
    __device__ __forceinline__ u64 king(int sq) const { return p[320 + sq]; }
    __device__ __forceinline__ u64 pawn(int black, int sq, int k) const { return p[384 + black * 128 + sq * 2 + k]; }
    // word i of the shared copy, from the device-memory tables (the kernel's threads fill the copy cooperatively)
    static __device__ __forceinline__ u64 source_word(int i) {
        if (i < 256) return __ldg(&g_geom_dev.line[0][0] + i);
        if (i < 320) return __ldg(&g_geom_dev.knight[0] + (i - 256));
        if (i < 384) return __ldg(&g_geom_dev.king[0] + (i - 320));
        return __ldg(&g_geom_dev.pawn[0][0][0] + (i - 384));
    }
};
#endif
#define GCB_RAY_DESC_MASK 0x35u  // entries 0,2,4,5: nearest square first = highest bit first
GCB_HD int order_class(int code, int white) {
    return code == PC_KING ? 1 : code == PC_KNIGHT ? 2 : code == PC_PAWN ? (white ? 3 : 4) : 0;
}

// ---- arithmetic line masks (include the square itself); only used on rare paths (pins, check masks)
GCB_HD u64 mask_file(int sq) { return GCB_FILE_A << (sq & 7); }
GCB_HD u64 mask_rank(int sq) { return 0xFFULL << (sq & 56); }
GCB_HD u64 mask_diag(int sq) {  // steps of +-9: (row-col) constant
    int d = (sq >> 3) - (sq & 7);
    return d >= 0 ? (GCB_MAIN_DIAG << (8 * d)) : (GCB_MAIN_DIAG >> (-8 * d));
}
GCB_HD u64 mask_anti(int sq) {  // steps of +-7: (row+col) constant
    int s = (sq >> 3) + (sq & 7) - 7;
    return s >= 0 ? (GCB_ANTI_DIAG << (8 * s)) : (GCB_ANTI_DIAG >> (-8 * s));
}

// Hyperbola Quintessence on one line.  maskEx excludes the square.  The result holds every
// square of the line up to AND INCLUDING the first occupied one in both directions -- exactly
// the reference's ray rule "empty -> add, continue; any piece -> add, stop" in attack mode
// (lib.rs:1089-1104); play mode masks own pieces afterwards (lib.rs:1063-1081).
GCB_HD u64 hq_line(u64 occ, u64 maskEx, u64 bit, u64 rbit) {
    u64 o = occ & maskEx;
    u64 f = o - bit;
    u64 r = gcb_brev64(gcb_brev64(o) - rbit);
    return (f ^ r) & maskEx;
}
template <class G>
GCB_HD u64 rook_att(const G& geo, int sq, u64 occ) {
    u64 bit = 1ULL << sq, rbit = 1ULL << (63 - sq);
    return hq_line(occ, geo.line(sq, 0), bit, rbit) | hq_line(occ, geo.line(sq, 1), bit, rbit);
}
template <class G>
GCB_HD u64 bishop_att(const G& geo, int sq, u64 occ) {
    u64 bit = 1ULL << sq, rbit = 1ULL << (63 - sq);
    return hq_line(occ, geo.line(sq, 2), bit, rbit) | hq_line(occ, geo.line(sq, 3), bit, rbit);
}
GCB_HD u64 rook_att(int sq, u64 occ) { return rook_att(GeomGlobal(), sq, occ); }
GCB_HD u64 bishop_att(int sq, u64 occ) { return bishop_att(GeomGlobal(), sq, occ); }

// ---- set-wise leaper attacks (all pieces of the set at once)
GCB_HD u64 knight_set_att(u64 n) {
    u64 l1 = (n >> 1) & ~GCB_FILE_H, l2 = (n >> 2) & ~(GCB_FILE_G | GCB_FILE_H);
    u64 r1 = (n << 1) & ~GCB_FILE_A, r2 = (n << 2) & ~(GCB_FILE_A | GCB_FILE_B);
    u64 h1 = l1 | r1, h2 = l2 | r2;
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
GCB_HD u64 king_set_att(u64 k) {
    u64 a = ((k << 1) & ~GCB_FILE_A) | ((k >> 1) & ~GCB_FILE_H);
    u64 row = a | k;
    return a | (row << 8) | (row >> 8);
}
// pawn diagonals of a whole pawn set; white pawns attack towards row-1 (lib.rs:920-924)
GCB_HD u64 pawn_set_att(u64 p, int white) {
    if (white) return ((p & ~GCB_FILE_H) >> 7) | ((p & ~GCB_FILE_A) >> 9);
    return ((p & ~GCB_FILE_H) << 9) | ((p & ~GCB_FILE_A) << 7);
}

// ---- the king square the reference's check test looks at (lib.rs:644-653): the inner-loop
// `break` means: LAST row that holds a king of that colour, FIRST column within it (Q15).
GCB_HD int ref_king_square(u64 kings) {
    int top = gcb_msb(kings);     // highest set bit -> its row is the last row
    u64 rowbits = kings & (0xFFULL << (top & 56));
    return gcb_lsb(rowbits);
}

// Is `sq` in the attack/defence map of `side` (lib.rs:669-677)?  Symmetric form: a slider attacks sq
// iff sq's own ray reaches it, leapers likewise; used where only one square matters.  (The Q14
// pawn exclusion cannot apply: callers pass the square of a king of the OTHER colour.)
GCB_HD bool square_attacked_by(const Board& b, int sq, u64 side, int side_is_white) {
    const u64 bit = 1ULL << sq, occ = bb_occ(b);
    u64 leap = (knight_set_att(bit) & bb_knights(b)) | (king_set_att(bit) & bb_kings(b)) |
               (pawn_set_att(bit, !side_is_white) & bb_pawns(b));
    if (leap & side) return true;
    u64 q = bb_queens(b);
    if (rook_att(sq, occ) & (bb_rooks(b) | q) & side) return true;
    if (bishop_att(sq, occ) & (bb_bishops(b) | q) & side) return true;
    return false;
}

// both check flags, lib.rs:1386-1393 (update_state): bit0 = white king checked, bit1 = black
GCB_HD u32 check_flags(const Board& b) {
    u64 occ = bb_occ(b), white = b.w, black = occ & ~b.w, kings = bb_kings(b);
    u32 f = 0;
    u64 wk = kings & white, bk = kings & black;
    if (wk && square_attacked_by(b, ref_king_square(wk), black, 0)) f |= 1u;
    if (bk && square_attacked_by(b, ref_king_square(bk), white, 1)) f |= 2u;
    return f;
}

// State::new, lib.rs:315-322: castle flags are dropped when that side has no king on the board
GCB_HD u32 mask_rights(const Board& b, u32 rights) {
    u64 kings = bb_kings(b);
    if (!(kings & b.w)) rights &= ~(RT_WK | RT_WQ);
    if (!(kings & ~b.w)) rights &= ~(RT_BK | RT_BQ);
    return rights;
}

// ---------------------------------------------------------------------------------------------
// Legal move generation (lib.rs:460-610), B200 formulation.
//
// The reference simulates every candidate move and recomputes the whole opponent attack map
// (move_leaves_king_checked, lib.rs:612-626).  Here the legal set of a position is produced as ONE
// 64-bit target set per own piece ("piece slots", own pieces in ascending square order = the
// reference's row-major scan, lib.rs:510-511) plus two castle bits:
//   gen_prepare()  -- per position: opponent attack map, checkers of the (reference-rule, Q15) king
//                     square, the check mask (capture the single checker / block its ray) and the
//                     set of pinned pieces with their pin rays.  Exactly equivalent to simulate-and-
//                     recompute for NON-king moves because there is no en passant and a non-king move
//                     changes the attack on the king square only through (a) capturing an attacker,
//                     (b) blocking a ray, (c) vacating a ray through the king square.
//   gen_targets()  -- per own piece, TYPE-MAJOR (all rooks, then bishops, queens, knights, kings,
//                     pawns) so that the 32 envs of a warp execute the same code; the slot index is
//                     the piece's rank among the own pieces, so slots come out in reference order.
//   emit_piece_moves() / nth_target() -- the reference's per-piece move ORDER (direction order of
//                     lib.rs:797-806, 824-851, 891-900, 935-959; rays by increasing distance) is a
//                     pure function of (piece type, colour, square, target set): the ordered list
//                     is decoded from the slots only where a caller asks for it.
// ATTACK lists (get_possible_moves(attack=True)) keep a direct ordered generator (gen_attack_moves).
// ---------------------------------------------------------------------------------------------
GCB_HD u64 sq_bit_safe(int sq) { return (unsigned)sq < 64u ? (1ULL << sq) : 0ULL; }

// the queen line through both squares (0 when they are not aligned); a != b
GCB_HD u64 line_through(int a, int b) {
    const int ra = a >> 3, ca = a & 7, rb = b >> 3, cb = b & 7;
    u64 m = 0;
    if (ca == cb) m = mask_file(a);
    else if (ra == rb) m = mask_rank(a);
    else if (ra - ca == rb - cb) m = mask_diag(a);
    else if (ra + ca == rb + cb) m = mask_anti(a);
    return m;
}

struct GenCtx {
    u64 occ, own, enemy;  // the per-type sets are recomputed from the planes where they are used (2 LOP3 each): fewer live registers
    u64 eatt;     // opponent attack map on the CURRENT board, own king left on it (lib.rs:466-470; Q6); exact on the
                  // own kings' neighbourhoods and the castle squares, the only places it is read (see gen_prepare)
    u64 satt;     // attack map of the side to move, accumulated by gen_targets (check flag of the other side)
    u64 cm;       // targets that resolve the check for a non-king piece (all ones when not in check / no king)
    u64 pinned;   // own pieces whose removal opens a slider line onto the king square
    u64 pinrays;  // union of (between(king, pinner) | pinner)
    int ksq;
    bool has_king, in_check;
    int white;
};

template <class G>
GCB_HD void gen_prepare(const Board& b, int white_to_move, GenCtx& g, const G& geo) {
    g.white = white_to_move;
    g.occ = bb_occ(b);
    g.own = white_to_move ? b.w : (g.occ & ~b.w);
    g.enemy = g.occ & ~g.own;
    const u64 occ = g.occ, enemy = g.enemy;
    const u64 ekings = bb_kings(b) & enemy;
    const u64 eRQ = (bb_rooks(b) | bb_queens(b)) & enemy, eBQ = (bb_bishops(b) | bb_queens(b)) & enemy;

    // opponent attack map (lib.rs:669-677): pawns minus squares holding the attacker's OWN king (Q14)
    u64 eatt = pawn_set_att(bb_pawns(b) & enemy, !white_to_move) & ~ekings;
    eatt |= knight_set_att(bb_knights(b) & enemy) | king_set_att(ekings);
    const u64 ownk = bb_kings(b) & g.own;
    u64 zR = eRQ, zB = eBQ;
#if GCB_ZONE_FILTER
    // The map is only ever tested on the own kings' neighbourhoods (king targets) and on the castle squares (gen_castles),
    // so a slider none of whose lines crosses that zone need not be evaluated: with one own king, two table words pick the
    // sliders that matter (fewer trips of the two loops below for the whole warp).  Several own kings (Q15), or black to
    // move with a WHITE king on e8 (the castle test of Q3 then looks at c8..g8): no filter.  No own king: nothing reads it.
    if (!ownk) zR = zB = 0;
    else if (!(ownk & (ownk - 1)) && (white_to_move || !((bb_kings(b) & b.w) >> 4 & 1ULL))) {
        const int k = gcb_msb(ownk);
        zR &= GCB_GEOM(zone[0][k]), zB &= GCB_GEOM(zone[1][k]);
    }
#endif
    for (u64 s = zR; s;) eatt |= rook_att(geo, gcb_take(s), occ);
    for (u64 s = zB; s;) eatt |= bishop_att(geo, gcb_take(s), occ);
    g.eatt = eatt;

    g.satt = 0, g.cm = ~0ULL, g.pinned = 0, g.pinrays = 0, g.ksq = 0, g.in_check = false;
    g.has_king = ownk != 0;
    if (!g.has_king) return;  // lib.rs:655-658: no king -> nothing is filtered
    const int ksq = ref_king_square(ownk);
    const u64 kbit = 1ULL << ksq;
    g.ksq = ksq;
    // pieces that attack the king square right now (attack sets are symmetric; an enemy pawn attacks ksq iff it
    // stands where a pawn of the MOVER's colour on ksq would attack; enemy kings count, Q22)
    u64 chk = (geo.knight(ksq) & bb_knights(b) & enemy) | (geo.king(ksq) & ekings) |
              (geo.pawn(!white_to_move, ksq, 1) & bb_pawns(b) & enemy);
    // enemy sliders on a line through the king square that they move along: nothing in between -> checker; exactly
    // one piece in between and it is ours -> that piece is pinned to the ray (between | pinner)
    const u64 cand = (eRQ & (geo.line(ksq, 0) | geo.line(ksq, 1))) | (eBQ & (geo.line(ksq, 2) | geo.line(ksq, 3)));
    for (u64 p = cand; p;) {
        const int psq = gcb_take(p);
        const u64 btw = GCB_GEOM(between[ksq][psq]), blockers = btw & occ;
        if (!blockers) chk |= 1ULL << psq;
        else if (!(blockers & (blockers - 1)) && (blockers & g.own)) g.pinned |= blockers, g.pinrays |= btw | (1ULL << psq);
    }
    g.in_check = chk != 0;
    if (chk) {
        if (chk & (chk - 1)) g.cm = 0;  // two or more attackers: no non-king move can remove both
        else g.cm = chk | GCB_GEOM(between[ksq][gcb_lsb(chk)]);
    }
    (void)kbit;
}

GCB_HD void gen_prepare(const Board& b, int white_to_move, GenCtx& g) { gen_prepare(b, white_to_move, g, GeomGlobal()); }

// the squares a pinned piece on `sq` may move to: its own pin ray (between(king, pinner) | pinner)
GCB_HD u64 pin_mask(const GenCtx& g, int sq) {
    const u64 kbit = 1ULL << g.ksq;
    const u64 side = sq > g.ksq ? ~(kbit | (kbit - 1)) : (kbit - 1);
    return g.pinrays & line_through(g.ksq, sq) & side;
}

// Targets of every own piece in `subset` (a set of squares; the caller passes all own pieces, or a
// chunk of them when there are more pieces than slots).  sink.put(rank, targets): rank = index of the
// piece among the own pieces of `subset` in ascending square order.
template <class Sink, class G>
GCB_HD void gen_targets(const Board& b, GenCtx& g, u64 subset, Sink& sink, const G& geo) {
    const u64 occ = g.occ, own = g.own, notown = ~g.own;
    const u64 mine = own & subset;
#define GCB_PUT(sq_, below_, T_)                                       \
    do {                                                               \
        const u64 t__ = (T_);                                          \
        sink.put(gcb_popc(mine & (below_)), t__);                      \
    } while (0)
    // rooks
    for (u64 s = bb_rooks(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        const u64 a = rook_att(geo, sq, occ);
        g.satt |= a;
        GCB_PUT(sq, below, a & notown & g.cm);
    }
    // bishops
    for (u64 s = bb_bishops(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        const u64 a = bishop_att(geo, sq, occ);
        g.satt |= a;
        GCB_PUT(sq, below, a & notown & g.cm);
    }
    // queens
    for (u64 s = bb_queens(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        const u64 a = rook_att(geo, sq, occ) | bishop_att(geo, sq, occ);
        g.satt |= a;
        GCB_PUT(sq, below, a & notown & g.cm);
    }
    // knights
    for (u64 s = bb_knights(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        const u64 a = geo.knight(sq);
        g.satt |= a;
        GCB_PUT(sq, below, a & notown & g.cm);
    }
    // kings: never passed through the legality filter (lib.rs:615-619); attack map with the king on it (Q6)
    for (u64 s = bb_kings(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        const u64 a = geo.king(sq);
        g.satt |= a;
        GCB_PUT(sq, below, a & notown & ~g.eatt);
    }
    // pawns (lib.rs:918-964): one step if empty; two steps from the start row if the TARGET is empty (the
    // jumped square is not tested, Q13); diagonals onto enemy pieces incl. the king; no en passant
    {
        const u64 pw = bb_pawns(b) & mine, allp = bb_pawns(b) & own;
        g.satt |= pawn_set_att(pw, g.white) & ~(bb_kings(b) & own);  // Q14
        (void)allp;
        const u64 free_cm = ~occ & g.cm, enemy_cm = g.enemy & g.cm;  // (hoisted: two LOP3 per half and pawn instead of three)
        for (u64 s = pw; s;) {
            u64 below;
            const int sq = gcb_take(s, below);
            const u64 push = geo.pawn(!g.white, sq, 0), cap = geo.pawn(!g.white, sq, 1);
            GCB_PUT(sq, below, (push & free_cm) | (cap & enemy_cm));
        }
    }
#undef GCB_PUT
    // pinned pieces (rare per position, but some env of a warp nearly always has one): one fix-up pass over the
    // slots instead of a pin test at every generation site
    for (u64 p = g.pinned & mine & ~bb_kings(b); p;) {
        u64 below;
        const int sq = gcb_take(p, below), r = gcb_popc(mine & below);
        const u64 t = sink.get(r), t2 = t & pin_mask(g, sq);
        sink.replace(r, t, t2);
    }
}

template <class Sink>
GCB_HD void gen_targets(const Board& b, GenCtx& g, u64 subset, Sink& sink) {
    gen_targets(b, g, subset, sink, GeomGlobal());
}

// castles, lib.rs:578-610 + 966-1056: needs the mover's king on the board and K-right OR Q-right (Q4).
// The black branch tests WHITE ids (+ROOK on a8/h8, +KING on e8) exactly like lib.rs:1023-1046 (Q3) -- never
// true in play.  bit0 = queen side (listed first, lib.rs:992), bit1 = king side.
GCB_HD u32 gen_castles(const Board& b, const GenCtx& g, u32 rights) {
    if (!g.has_king) return 0;
    const u64 wR = bb_rooks(b) & b.w, wK = bb_kings(b) & b.w, occ = g.occ, eatt = g.eatt;
    u32 c = 0;
    if (g.white) {
        if (!(rights & (RT_WK | RT_WQ))) return 0;
        const u64 e1 = 1ULL << 60;
        if ((wR >> 56 & 1) && !(occ & (7ULL << 57)) && (wK & e1) && !(eatt & (7ULL << 58))) c |= 1u;
        if ((wR >> 63 & 1) && !(occ & (3ULL << 61)) && (wK & e1) && !(eatt & (7ULL << 60))) c |= 2u;
    } else {
        if (!(rights & (RT_BK | RT_BQ))) return 0;
        const u64 e8 = 1ULL << 4;
        if ((wR & 1) && !(occ & (7ULL << 1)) && (wK & e8) && !(eatt & (7ULL << 2))) c |= 1u;
        if ((wR >> 7 & 1) && !(occ & (3ULL << 5)) && (wK & e8) && !(eatt & (7ULL << 4))) c |= 2u;
    }
    return c;
}
GCB_HD int castle_action(int white, int king_side) {
    return white ? (king_side ? ACT_CASTLE_KS_WHITE : ACT_CASTLE_QS_WHITE) : (king_side ? ACT_CASTLE_KS_BLACK : ACT_CASTLE_QS_BLACK);
}

// ---- the reference's move ORDER inside one piece: a walk over the 8 ordered direction masks of its class
template <class Emit>
GCB_HD void emit_piece_moves(Emit& em, int code, int white, int sq, u64 T) {
    if (!T) return;
    const int base = sq * 64, cls = order_class(code, white);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 8; k++) {
        u64 m = T & GCB_GEOM(ord[cls][sq][k]);
        if ((GCB_RAY_DESC_MASK >> k) & 1) {
            while (m) {
                const int to = gcb_msb(m);
                m ^= 1ULL << to;
                em.push(base + to);
            }
        } else {
            while (m) {
                const int to = gcb_lsb(m);
                m &= m - 1;
                em.push(base + to);
            }
        }
    }
}

// the idx-th (0-based) target of one piece in the reference's order; idx < popc(T).  Same instructions for every
// piece type: find the direction that holds the idx-th target, then walk idx squares along it (nearest first).
// ROLLED: a loop with early exit (8x less code: what the multi-step kernel wants, whose hot loop has to fit the instruction
// cache) or all 8 table loads issued at once (what the single-step kernels want: their caches are cold at every launch and
// the rolled loop's dependent load -> popcount -> branch chain is paid at L2 latency)
template <bool ROLLED = (GCB_ROLL_NTH != 0)>
GCB_HD int nth_target(int code, int white, int sq, u64 T, int idx) {
    const int cls = order_class(code, white);
    u64 mf = 0;
    bool desc = false;
#if GCB_NTH_BSEARCH
    // the direction that holds the idx-th target = the first k with popc(T & cum[k]) > idx: a search over the cumulative
    // direction masks, the same instructions in every lane (the walk over the directions ran the warp for its slowest lane:
    // 10 of 32 lanes active).  `below` / `above` end up as cum[k-1] / cum[k], so the direction's own targets need no
    // further load.
    {
        (void)ROLLED;
        int k = 0, nb = 0;
        u64 below = 0, above = ~0ULL;
#if GCB_NTH_TWO_LEVEL
        {  // two round trips instead of three: the quarter from cum[1], cum[3], cum[5] (issued together), then one more word
            const u64 m1 = GCB_GEOM(cum[cls][sq][1]), m3 = GCB_GEOM(cum[cls][sq][3]), m5 = GCB_GEOM(cum[cls][sq][5]);
            const int c1 = gcb_popc(T & m1), c3 = gcb_popc(T & m3), c5 = gcb_popc(T & m5);
            const bool r1 = idx >= c1, r3 = idx >= c3, r5 = idx >= c5;  // (monotone: r5 implies r3 implies r1)
            k = r5 ? 6 : r3 ? 4 : r1 ? 2 : 0, nb = r5 ? c5 : r3 ? c3 : r1 ? c1 : 0;
            below = r5 ? m5 : r3 ? m3 : r1 ? m1 : 0ULL, above = r5 ? ~0ULL : r3 ? m5 : r1 ? m3 : m1;
            const u64 m = GCB_GEOM(cum[cls][sq][k]);
            const int c = gcb_popc(T & m);
            const bool right = idx >= c;
            k = right ? k + 1 : k, nb = right ? c : nb;
            below = right ? m : below, above = right ? above : m;
        }
#else
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int step = 4; step; step >>= 1) {
            const u64 m = GCB_GEOM(cum[cls][sq][k + step - 1]);
            const int c = gcb_popc(T & m);
            const bool right = idx >= c;
            k = right ? k + step : k, nb = right ? c : nb;
            below = right ? m : below, above = right ? above : m;
        }
#endif
        mf = T & above & ~below, idx -= nb;
        desc = (GCB_RAY_DESC_MASK >> k) & 1;
    }
#else
  if (ROLLED) {
    // a rolled loop with early exit: 8x less code on the step kernel's hot path (its instruction footprint is what the
    // instruction cache holds)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < 8; k++) {
        const u64 m = T & GCB_GEOM(ord[cls][sq][k]);
        const int c = gcb_popc(m);
        if (idx < c) {
            mf = m, desc = (GCB_RAY_DESC_MASK >> k) & 1;
            break;
        }
        idx -= c;
    }
  } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 8; k++) {
        const u64 m = T & GCB_GEOM(ord[cls][sq][k]);
        const int c = gcb_popc(m);
        const bool hit = !mf && idx < c;
        if (hit) mf = m, desc = (GCB_RAY_DESC_MASK >> k) & 1;
        if (!mf) idx -= c;
    }
  }
#endif
    if (!mf) return sq;
    if (desc) mf = gcb_brev64(mf);  // nearest first = lowest bit first in both cases
    const int t = gcb_select64(mf, idx);
    return desc ? 63 - t : t;
}

// ---- attack=True lists (get_possible_moves(attack=True), lib.rs:928-933, 1089-1104, 1147-1174): defended
// squares included, pawn diagonals only (minus squares holding the mover's own king, Q14), no legality filter,
// no castles; ordered, straight from the board.
template <class Emit>
GCB_HD void gen_attack_moves(const Board& b, int white_to_move, Emit& em) {
    const u64 occ = bb_occ(b);
    const u64 own = white_to_move ? b.w : (occ & ~b.w);
    const u64 ownk = bb_kings(b) & own;
    for (u64 todo = own; todo; todo &= todo - 1) {
        const int sq = gcb_lsb(todo);
        const u64 bit = 1ULL << sq;
        const int code = piece_code(b, sq);
        u64 T;
        if (code == PC_PAWN) T = pawn_set_att(bit, white_to_move) & ~ownk;
        else if (code == PC_KNIGHT) T = knight_set_att(bit);
        else if (code == PC_KING) T = king_set_att(bit);
        else {
            T = 0;
            if (code == PC_ROOK || code == PC_QUEEN) T |= rook_att(sq, occ);
            if (code == PC_BISHOP || code == PC_QUEEN) T |= bishop_att(sq, occ);
        }
        if (code == PC_PAWN) {  // (row-p, col+1) then (row-p, col-1)
            const int d1 = white_to_move ? -7 : 9, d2 = white_to_move ? -9 : 7;
            if (T & sq_bit_safe(sq + d1)) em.push(sq * 64 + sq + d1);
            if (T & sq_bit_safe(sq + d2)) em.push(sq * 64 + sq + d2);
        } else {
            emit_piece_moves(em, code, white_to_move, sq, T);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Ordered list from the slots, TYPE-MAJOR.  The position of a move in the reference-ordered list is
//     offset(piece) + rank of the move inside the piece,     offset = prefix sum of the slot popcounts,
// so once the offsets are known the pieces can be decoded in ANY order: all rooks, then bishops, ... -- the 32
// positions of a warp run the same code (a rook's four rays, a knight's eight jumps) instead of 32 different
// piece types.  Slots: get(r).  Offs: set(r, v) / get(r) (16 small integers of scratch).  Out: put(pos, action).
// `mine` = the own pieces of this chunk (<= GCB_SLOTS), `base` = list length before the chunk; returns the new length.
// ---------------------------------------------------------------------------------------------
template <class Out>
GCB_HD void emit_ray(Out& out, int& pos, int from64, u64 m, bool desc) {
    if (desc) {
        while (m) {
            const int to = gcb_msb(m);
            m ^= 1ULL << to;
            out.put(pos++, from64 + to);
        }
    } else {
        while (m) {
            const int to = gcb_lsb(m);
            m &= m - 1;
            out.put(pos++, from64 + to);
        }
    }
}

template <class Slots, class Offs, class Out>
GCB_HD int emit_chunk_typemajor(const Board& b, int white, u64 mine, const Slots& slots, Offs& offs, Out& out, int base) {
    {
        const int np = gcb_popc(mine);
        int acc = base;
        for (int r = 0; r < np; r++) {
            offs.set(r, acc);
            acc += gcb_popc(slots.get(r));
        }
        base = acc;
    }
    const u64 t0 = b.t0, t1 = b.t1, t2 = b.t2;
#define GCB_PIECE(set_)                                             \
    u64 below;                                                      \
    const int sq = gcb_take(set_, below);                           \
    const int r = gcb_popc(mine & below);                           \
    const u64 T = slots.get(r);                                     \
    int pos = offs.get(r);                                          \
    const int from64 = sq * 64
    // rooks: rays 0-3
    for (u64 s = (t0 & t1 & ~t2) & mine; s;) {
        GCB_PIECE(s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) emit_ray(out, pos, from64, T & GCB_GEOM(ord[0][sq][k]), (GCB_RAY_DESC_MASK >> k) & 1);
    }
    // bishops: rays 4-7
    for (u64 s = (~t0 & ~t1 & t2) & mine; s;) {
        GCB_PIECE(s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 4; k < 8; k++) emit_ray(out, pos, from64, T & GCB_GEOM(ord[0][sq][k]), (GCB_RAY_DESC_MASK >> k) & 1);
    }
    // queens: all eight
    for (u64 s = (~t0 & t1 & ~t2) & mine; s;) {
        GCB_PIECE(s);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 8; k++) emit_ray(out, pos, from64, T & GCB_GEOM(ord[0][sq][k]), (GCB_RAY_DESC_MASK >> k) & 1);
    }
    // knights (lib.rs:891-900) and kings (lib.rs:797-806): eight single squares in the reference's order
    for (u64 s = (t0 & ~t1 & t2) & mine; s;) {
        GCB_PIECE(s);
        const int d[8] = {-17, -15, 15, 17, -10, -6, 6, 10};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 8; k++)
            if (T & sq_bit_safe(sq + d[k])) out.put(pos++, from64 + sq + d[k]);
    }
    for (u64 s = (t0 & ~t1 & ~t2) & mine; s;) {
        GCB_PIECE(s);
        const int d[8] = {8, -8, 1, -1, 9, 7, -7, -9};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 8; k++)
            if (T & sq_bit_safe(sq + d[k])) out.put(pos++, from64 + sq + d[k]);
    }
    // pawns (lib.rs:935-959): one step, two steps, (row-p, col+1), (row-p, col-1)
    {
        const int d0 = white ? -8 : 8, d1 = white ? -16 : 16, d2 = white ? -7 : 9, d3 = white ? -9 : 7;
        for (u64 s = (~t0 & t1 & t2) & mine; s;) {
            GCB_PIECE(s);
            if (T & sq_bit_safe(sq + d0)) out.put(pos++, from64 + sq + d0);
            if (T & sq_bit_safe(sq + d1)) out.put(pos++, from64 + sq + d1);
            if (T & sq_bit_safe(sq + d2)) out.put(pos++, from64 + sq + d2);
            if (T & sq_bit_safe(sq + d3)) out.put(pos++, from64 + sq + d3);
        }
    }
#undef GCB_PIECE
    return base;
}

// Whole ordered legal list of one position through a small slot buffer (any number of own pieces): the
// engine-level get_possible_moves.  Returns the list length.
#define GCB_SLOTS 16
template <class Slots, class Offs, class Out>
GCB_HD int gen_legal_list(const Board& b, int white_to_move, u32 rights, Slots& slots, Offs& offs, Out& out, bool* in_check_out) {
    GenCtx g;
    gen_prepare(b, white_to_move, g);
    if (in_check_out) *in_check_out = g.in_check;
    int n = 0;
    u64 rem = g.own;
    while (rem) {
        // next chunk of at most GCB_SLOTS own pieces in square order
        u64 chunk = rem;
        if (gcb_popc(rem) > GCB_SLOTS) {
            u64 t = rem;
            for (int i = 0; i < GCB_SLOTS; i++) t &= t - 1;
            chunk = rem ^ t;
        }
        rem ^= chunk;
        gen_targets(b, g, chunk, slots);
        n = emit_chunk_typemajor(b, white_to_move, chunk, slots, offs, out, n);
    }
    const u32 c = gen_castles(b, g, rights);
    if (c & 1u) out.put(n++, castle_action(white_to_move, 0));
    if (c & 2u) out.put(n++, castle_action(white_to_move, 1));
    return n;
}

// ---- attack=True lists through the same two stages as the legal lists: per own piece its attack / defence set (rays up
// to and including the first piece of either colour, lib.rs:1089-1104; knight / king neighbourhoods, lib.rs:1147-1174;
// pawn diagonals minus squares holding the mover's own king, lib.rs:928-933, Q14), type-major, then the ordered decode.
// No legality filter, no castles.  (gen_attack_moves above is the direct square-major form; the host-compiled tests use both.)
template <class Sink, class G>
GCB_HD void gen_attack_targets(const Board& b, int white, u64 subset, Sink& sink, const G& geo) {
    const u64 occ = bb_occ(b), own = white ? b.w : (occ & ~b.w), mine = own & subset, ownk = bb_kings(b) & own;
#define GCB_APUT(below_, T_) sink.put(gcb_popc(mine & (below_)), (T_))
    for (u64 s = bb_rooks(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        GCB_APUT(below, rook_att(geo, sq, occ));
    }
    for (u64 s = bb_bishops(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        GCB_APUT(below, bishop_att(geo, sq, occ));
    }
    for (u64 s = bb_queens(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        GCB_APUT(below, rook_att(geo, sq, occ) | bishop_att(geo, sq, occ));
    }
    for (u64 s = bb_knights(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        GCB_APUT(below, geo.knight(sq));
    }
    for (u64 s = bb_kings(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        GCB_APUT(below, geo.king(sq));
    }
    for (u64 s = bb_pawns(b) & mine; s;) {
        u64 below;
        const int sq = gcb_take(s, below);
        GCB_APUT(below, geo.pawn(!white, sq, 1) & ~ownk);
    }
#undef GCB_APUT
}

template <class Slots, class Offs, class Out>
GCB_HD int gen_attack_list(const Board& b, int white_to_move, Slots& slots, Offs& offs, Out& out) {
    const u64 occ = bb_occ(b);
    int n = 0;
    u64 rem = white_to_move ? b.w : (occ & ~b.w);
    while (rem) {
        u64 chunk = rem;
        if (gcb_popc(rem) > GCB_SLOTS) {
            u64 t = rem;
            for (int i = 0; i < GCB_SLOTS; i++) t &= t - 1;
            chunk = rem ^ t;
        }
        rem ^= chunk;
        gen_attack_targets(b, white_to_move, chunk, slots, GeomGlobal());
        n = emit_chunk_typemajor(b, white_to_move, chunk, slots, offs, out, n);
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// Zobrist key of the board only (the reference's repetition key is the 64-char board string,
// chess_v2.py:404-407, 599-602: no side to move, no rights).  Keys are splitmix64 of
// (piece index, square); piece index = id + 6 in 0..12.  The env kernels read them from a
// 13*64-entry table in global memory (fill_zobrist_table), three L1-resident loads per ply.
// ---------------------------------------------------------------------------------------------
GCB_HD u64 gcb_splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
#define GCB_ZOB_ENTRIES (13 * 64)
GCB_HD u64 zobrist_piece(int pid /* signed id != 0 */, int sq) {
    return gcb_splitmix64((u64)((pid + 6) * 64 + sq) + 0x6A09E667F3BCC908ULL);
}
GCB_HD void fill_zobrist_entry(u64* tab, int i) { tab[i] = (i >> 6) == 6 ? 0ULL : zobrist_piece((i >> 6) - 6, i & 63); }
GCB_HD u64 zob_at(const u64* tab, int pid, int sq) {
#if defined(__CUDA_ARCH__)
    return __ldg(tab + (pid + 6) * 64 + sq);
#else
    return tab[(pid + 6) * 64 + sq];
#endif
}
GCB_HD u64 zobrist_full(const Board& b) {
    u64 occ = bb_occ(b), k = 0;
    while (occ) {
        int sq = gcb_lsb(occ);
        occ &= occ - 1;
        k ^= zobrist_piece(piece_id(b, sq), sq);
    }
    return k;
}
// what the history ring stores / compares: never 0, which marks "no ply in this slot"
GCB_HD u64 hist_key(u64 zkey) { return zkey | 1ULL; }

// ---------------------------------------------------------------------------------------------
// next_state (lib.rs:679-784) on planes.  `rights` must already be masked (mask_rights).
// Returns the reward; *status = 0 ok, -1 empty from-square (reference panics), -2 bad action.
// *irreversible is set for pawn moves and captures (board-only repetition key can never recur).
// ztab != NULL: *zk (Zobrist key of the board) is updated incrementally.
// ---------------------------------------------------------------------------------------------
GCB_HD int piece_value(int code) {
    // P1 N3 B3 R5 Q10 K0 (lib.rs:19-25), indexed by code 0..7, 4 bits each
    return (int)((0x01335A00u >> (code * 4)) & 15u);
}
GCB_HD void set_code(Board& b, u64 bit, int code, int white) {  // the square must be clear
    if (code & 1) b.t0 |= bit;
    if (code & 2) b.t1 |= bit;
    if (code & 4) b.t2 |= bit;
    if (white && code) b.w |= bit;
}

GCB_HD int apply_action(Board& b, u32& rights, int white_to_move, int action, int* status, bool* irreversible,
                        u64* zk = nullptr, const u64* ztab = nullptr) {
    *status = 0;
    *irreversible = false;
    int reward = 0;
    if (action < 4096) {
        if (action < 0) { *status = -2; return 0; }
        const int from = action >> 6, to = action & 63;
        const u64 fbit = 1ULL << from, tbit = 1ULL << to;
        const int code = piece_code(b, from);
        if (code == 0) { *status = -1; return 0; }
        const int fw = (int)((b.w >> from) & 1);
        const int cap = piece_code(b, to), capw = (int)((b.w >> to) & 1);
        int newcode = code, newwhite = fw;
        reward += piece_value(cap);
        // "pawn becomes queen" tests the WRONG ends (lib.rs:703-704, Q1); colour = the MOVER's
        if (code == PC_PAWN && ((white_to_move && (to >> 3) == 7) || (!white_to_move && (to >> 3) == 0))) {
            newcode = PC_QUEEN, newwhite = white_to_move;
            reward += 10;
        }
        clear_sq(b, fbit | tbit);
        set_code(b, tbit, newcode, newwhite);
        if (ztab) {
            u64 z = zob_at(ztab, fw ? code : -code, from) ^ zob_at(ztab, newwhite ? newcode : -newcode, to);
            if (cap) z ^= zob_at(ztab, capw ? cap : -cap, to);  // (from == to: the "captured" piece is the mover itself)
            *zk ^= z;
        }
        // rights react only to WHITE ids, column of the from-square only (lib.rs:711-734, Q5)
        if (fw && code == PC_KING) rights &= white_to_move ? ~(RT_WK | RT_WQ) : ~(RT_BK | RT_BQ);
        else if (fw && code == PC_ROOK) {
            if ((from & 7) == 0) rights &= white_to_move ? ~RT_WQ : ~RT_BQ;
            else if ((from & 7) == 7) rights &= white_to_move ? ~RT_WK : ~RT_BK;
        }
        *irreversible = (code == PC_PAWN) || (cap != 0);
    } else {
        const Board old = b;
        switch (action) {  // literal square writes, lib.rs:739-774
        case ACT_CASTLE_KS_WHITE:
            clear_sq(b, 15ULL << 60), set_code(b, 1ULL << 61, PC_ROOK, 1), set_code(b, 1ULL << 62, PC_KING, 1);
            rights &= ~(RT_WK | RT_WQ);
            break;
        case ACT_CASTLE_QS_WHITE:
            clear_sq(b, 31ULL << 56), set_code(b, 1ULL << 58, PC_KING, 1), set_code(b, 1ULL << 59, PC_ROOK, 1);
            rights &= ~(RT_WK | RT_WQ);
            break;
        case ACT_CASTLE_KS_BLACK:
            clear_sq(b, 15ULL << 4), set_code(b, 1ULL << 5, PC_ROOK, 0), set_code(b, 1ULL << 6, PC_KING, 0);
            rights &= ~(RT_BK | RT_BQ);
            break;
        case ACT_CASTLE_QS_BLACK:
            clear_sq(b, 31ULL), set_code(b, 1ULL << 2, PC_KING, 0), set_code(b, 1ULL << 3, PC_ROOK, 0);
            rights &= ~(RT_BK | RT_BQ);
            break;
        default: *status = -2; return 0;
        }
        if (ztab) {  // rare: key over the squares that changed
            u64 diff = (old.t0 ^ b.t0) | (old.t1 ^ b.t1) | (old.t2 ^ b.t2) | (old.w ^ b.w);
            for (; diff; diff &= diff - 1) {
                const int sq = gcb_lsb(diff);
                const int po = piece_id(old, sq), pn = piece_id(b, sq);
                if (po) *zk ^= zob_at(ztab, po, sq);
                if (pn) *zk ^= zob_at(ztab, pn, sq);
            }
        }
    }
    return reward;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  draw = word 0 of ctr=(env, episode, step, purpose),
// key=(seed lo, seed hi); identical to oracle/gc_oracle.c:gco_draw_u32.
// ---------------------------------------------------------------------------------------------
GCB_HD u32 philox_draw(u64 seed, u32 env, u32 episode, u32 step, u32 purpose) {
    u32 c0 = env, c1 = episode, c2 = step, c3 = purpose, k0 = (u32)seed, k1 = (u32)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u32 h0 = gcb_umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        u32 h1 = gcb_umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        u32 n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0, c1 = l1, c2 = n2, c3 = l0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    return c0;
}

// ---- mailbox <-> planes (wire format of the reference state dict: board int8[64])
GCB_HD Board board_from_mailbox(const int8_t* m) {
    Board b = {0, 0, 0, 0};
    for (int sq = 0; sq < 64; sq++) {
        int id = m[sq];
        int code = id < 0 ? -id : id;
        u64 bit = 1ULL << sq;
        if (code & 1) b.t0 |= bit;
        if (code & 2) b.t1 |= bit;
        if (code & 4) b.t2 |= bit;
        if (id > 0) b.w |= bit;
    }
    return b;
}
