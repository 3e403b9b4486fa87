// chess_core.cuh -- device-side rules of the gym-chess v2 env on bitboards (sm_100a).
//
// Re-design, not a translation: the reference walks an 8x8 isize mailbox with HashMap
// attack maps and simulates every candidate move (src/lib.rs:501-677); here a position is
// four 64-bit planes (three piece-code bit-planes + a colour plane, 32 B), slider attacks
// come from Hyperbola-Quintessence with __brevll on arithmetic line masks (no magic
// tables), the enemy attack map is one u64, and the legality filter is a symmetric
// "is the king square attacked after the move" test with a provably exact fast path.
// The OBSERVABLE behaviour (move order, quirks Q1-Q23 of SURVEY.md section 9) is that of
// the reference; each routine cites the lines whose results it reproduces.
//
// Square index = row*8 + col, row 0 = rank 8 (lib.rs:41-50, 1235-1238), so a row-major
// scan of the board (lib.rs:510-511) is an ascending bit scan.
#pragma once
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

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
GCB_HD int gcb_msb(u64 x) {  // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)x);
#else
    return 63 - __builtin_clzll(x);
#endif
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

// ---- arithmetic line masks (include the square itself)
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
GCB_HD u64 rook_att(int sq, u64 occ) {
    u64 bit = 1ULL << sq, rbit = 1ULL << (63 - sq);
    return hq_line(occ, mask_file(sq) ^ bit, bit, rbit) | hq_line(occ, mask_rank(sq) ^ bit, bit, rbit);
}
GCB_HD u64 bishop_att(int sq, u64 occ) {
    u64 bit = 1ULL << sq, rbit = 1ULL << (63 - sq);
    return hq_line(occ, mask_diag(sq) ^ bit, bit, rbit) | hq_line(occ, mask_anti(sq) ^ bit, bit, rbit);
}

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

// Attack/defence map of one side (lib.rs:669-677 with the attack-mode branches of the piece
// generators): sliders/knights: every on-board target up to the first piece inclusive; pawns:
// both forward diagonals unless the square holds that side's OWN king (lib.rs:928-933, Q14);
// king: all 8 neighbours (the map passed down is empty, lib.rs:670-671).
GCB_HD u64 side_attack_map(const Board& b, u64 side, int side_is_white) {
    u64 occ = bb_occ(b);
    u64 kings = bb_kings(b) & side;
    u64 att = pawn_set_att(bb_pawns(b) & side, side_is_white) & ~kings;
    att |= knight_set_att(bb_knights(b) & side);
    att |= king_set_att(kings);
    u64 rq = (bb_rooks(b) | bb_queens(b)) & side;
    while (rq) {
        int sq = gcb_lsb(rq);
        rq &= rq - 1;
        att |= rook_att(sq, occ);
    }
    u64 bq = (bb_bishops(b) | bb_queens(b)) & side;
    while (bq) {
        int sq = gcb_lsb(bq);
        bq &= bq - 1;
        att |= bishop_att(sq, occ);
    }
    return att;
}

// Is `sq` in the attack map of `side`?  Symmetric form of side_attack_map(): a slider attacks sq
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
// Ordered legal move generation (lib.rs:460-610).  Emit must provide  void push(int action).
// Returns nothing; the caller reads the count from its Emit.  Outputs:
//   *in_check : mover's (reference) king square is in the opponent attack map
// ATTACK = true reproduces get_possible_moves(attack=True): defended squares included, pawn
// diagonals only, no legality filter, no castles, king sees an empty attack map.
// ---------------------------------------------------------------------------------------------
struct KingSafety {
    u64 occ, eRQ, eBQ, leapers;  // leapers = enemy N/K/P that attack the king square right now
    u64 klines;                  // queen lines through the king square (incl. it)
    int ksq;
    bool has_king, in_check;
};

// Would the mover's king square be attacked after moving a NON-king piece from->to ?
// Equals lib.rs:612-626 (simulate with next_state, recompute the opponent attack map, look the
// king square up): attack sets are symmetric, the captured piece (if any) stops attacking.
GCB_HD bool king_attacked_after(const KingSafety& ks, u64 fbit, u64 tbit) {
    u64 keep = ~tbit;
    if (ks.leapers & keep) return true;
    u64 occ2 = (ks.occ & ~fbit) | tbit;
    if (rook_att(ks.ksq, occ2) & ks.eRQ & keep) return true;
    if (bishop_att(ks.ksq, occ2) & ks.eBQ & keep) return true;
    return false;
}

GCB_HD bool nonking_move_legal(const KingSafety& ks, u64 fbit, u64 tbit) {
    if (!ks.has_king) return true;  // lib.rs:655-658: no king -> nothing is filtered
    // exact fast path: king not attacked now and the moved piece is not on a line through the
    // king square -> vacating `from` cannot open a slider line, placing on `to` can only block,
    // a capture only removes attackers.
    if (!ks.in_check && !(ks.klines & fbit)) return true;
    return !king_attacked_after(ks, fbit, tbit);
}

template <bool ATTACK, class Emit>
GCB_HD void emit_targets_desc(Emit& em, const KingSafety& ks, int from, u64 fbit, u64 m) {
    while (m) {  // nearest first on a ray that runs towards lower square indices
        int to = gcb_msb(m);
        u64 tbit = 1ULL << to;
        m ^= tbit;
        if (ATTACK || nonking_move_legal(ks, fbit, tbit)) em.push(from * 64 + to);
    }
}
template <bool ATTACK, class Emit>
GCB_HD void emit_targets_asc(Emit& em, const KingSafety& ks, int from, u64 fbit, u64 m) {
    while (m) {  // nearest first on a ray that runs towards higher square indices
        int to = gcb_lsb(m);
        u64 tbit = 1ULL << to;
        m ^= tbit;
        if (ATTACK || nonking_move_legal(ks, fbit, tbit)) em.push(from * 64 + to);
    }
}

template <bool ATTACK, class Emit>
GCB_HD void gen_moves(const Board& b, int white_to_move, u32 rights, Emit& em, u64* eatt_out,
                                          bool* in_check_out) {
    const u64 occ = bb_occ(b);
    const u64 own = white_to_move ? b.w : (occ & ~b.w);
    const u64 enemy = occ & ~own;
    const u64 kings = bb_kings(b), queens = bb_queens(b), rooks = bb_rooks(b), bishops = bb_bishops(b),
              knights = bb_knights(b), pawns = bb_pawns(b);

    // opponent attack map on the CURRENT board, own king left on it (lib.rs:466-470; Q6)
    u64 eatt = 0;
    if (!ATTACK) eatt = side_attack_map(b, enemy, !white_to_move);

    KingSafety ks;
    ks.occ = occ;
    ks.has_king = (kings & own) != 0;
    ks.ksq = 0, ks.in_check = false, ks.klines = 0, ks.leapers = 0;
    ks.eRQ = (rooks | queens) & enemy, ks.eBQ = (bishops | queens) & enemy;
    if (!ATTACK && ks.has_king) {
        int ksq = ref_king_square(kings & own);
        u64 kbit = 1ULL << ksq;
        ks.ksq = ksq;
        ks.in_check = (eatt >> ksq) & 1;
        ks.klines = mask_file(ksq) | mask_rank(ksq) | mask_diag(ksq) | mask_anti(ksq);
        // enemy pawns that attack ksq stand where a pawn of the MOVER's colour on ksq would attack
        ks.leapers = (knight_set_att(kbit) & knights & enemy) | (king_set_att(kbit) & kings & enemy) |
                     (pawn_set_att(kbit, white_to_move) & pawns & enemy);
    }
    if (eatt_out) *eatt_out = eatt;
    if (in_check_out) *in_check_out = ks.in_check;

    // row-major scan of own pieces (lib.rs:510-553)
    u64 todo = own;
    while (todo) {
        const int sq = gcb_lsb(todo);
        const u64 bit = 1ULL << sq;
        todo &= todo - 1;
        const int code = piece_code(b, sq);
        if (code == PC_PAWN) {
            // lib.rs:918-964
            const int col = sq & 7, row = sq >> 3;
            if (ATTACK) {
                // (row-p, col+1) then (row-p, col-1), skipped when it holds the OWN king (Q14)
                if (white_to_move) {
                    if (row > 0 && col < 7 && !((kings & own) >> (sq - 7) & 1)) em.push(sq * 64 + sq - 7);
                    if (row > 0 && col > 0 && !((kings & own) >> (sq - 9) & 1)) em.push(sq * 64 + sq - 9);
                } else {
                    if (row < 7 && col < 7 && !((kings & own) >> (sq + 9) & 1)) em.push(sq * 64 + sq + 9);
                    if (row < 7 && col > 0 && !((kings & own) >> (sq + 7) & 1)) em.push(sq * 64 + sq + 7);
                }
            } else {
                int one, two, capr, capl;
                bool can_one, can_two;
                if (white_to_move) {
                    one = sq - 8, two = sq - 16, capr = sq - 7, capl = sq - 9;
                    can_one = row > 0, can_two = row == 6;
                } else {
                    one = sq + 8, two = sq + 16, capr = sq + 9, capl = sq + 7;
                    can_one = row < 7, can_two = row == 1;
                }
                if (can_one && !((occ >> one) & 1) && nonking_move_legal(ks, bit, 1ULL << one)) em.push(sq * 64 + one);
                // double step tests only the TARGET square (lib.rs:942-954, Q13)
                if (can_two && !((occ >> two) & 1) && nonking_move_legal(ks, bit, 1ULL << two)) em.push(sq * 64 + two);
                if (can_one && col < 7 && ((enemy >> capr) & 1) && nonking_move_legal(ks, bit, 1ULL << capr))
                    em.push(sq * 64 + capr);
                if (can_one && col > 0 && ((enemy >> capl) & 1) && nonking_move_legal(ks, bit, 1ULL << capl))
                    em.push(sq * 64 + capl);
            }
        } else if (code == PC_KNIGHT) {
            // lib.rs:889-916, order (-2,-1)(-2,1)(2,-1)(2,1)(-1,-2)(-1,2)(1,-2)(1,2)
            u64 t = knight_set_att(bit);
            if (!ATTACK) t &= ~own;
            const int d[8] = {-17, -15, 15, 17, -10, -6, 6, 10};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int to = sq + d[k];
                if (to >= 0 && to < 64 && ((t >> to) & 1)) {
                    if (ATTACK || nonking_move_legal(ks, bit, 1ULL << to)) em.push(sq * 64 + to);
                }
            }
        } else if (code == PC_KING) {
            // lib.rs:789-822 + 1113-1174, order (1,0)(-1,0)(0,1)(0,-1)(1,1)(1,-1)(-1,1)(-1,-1).
            // King moves are never passed through the legality filter (lib.rs:615-619).
            u64 t = king_set_att(bit);
            if (!ATTACK) t &= ~eatt & ~own;
            const int d[8] = {8, -8, 1, -1, 9, 7, -7, -9};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int to = sq + d[k];
                if (to >= 0 && to < 64 && ((t >> to) & 1)) em.push(sq * 64 + to);
            }
        } else {
            // sliders, lib.rs:824-887: rook dirs (-1,0)(1,0)(0,-1)(0,1), bishop dirs
            // (-1,-1)(-1,1)(1,-1)(1,1); queen = rook dirs then bishop dirs; increasing distance
            const u64 rbit = 1ULL << (63 - sq);
            const u64 below = bit - 1, above = ~(bit | below);
            const u64 keepmask = ATTACK ? ~0ULL : ~own;
            if (code == PC_ROOK || code == PC_QUEEN) {
                u64 f = hq_line(occ, mask_file(sq) ^ bit, bit, rbit) & keepmask;
                u64 r = hq_line(occ, mask_rank(sq) ^ bit, bit, rbit) & keepmask;
                emit_targets_desc<ATTACK>(em, ks, sq, bit, f & below);
                emit_targets_asc<ATTACK>(em, ks, sq, bit, f & above);
                emit_targets_desc<ATTACK>(em, ks, sq, bit, r & below);
                emit_targets_asc<ATTACK>(em, ks, sq, bit, r & above);
            }
            if (code == PC_BISHOP || code == PC_QUEEN) {
                u64 dg = hq_line(occ, mask_diag(sq) ^ bit, bit, rbit) & keepmask;
                u64 an = hq_line(occ, mask_anti(sq) ^ bit, bit, rbit) & keepmask;
                emit_targets_desc<ATTACK>(em, ks, sq, bit, dg & below);  // (-1,-1)
                emit_targets_desc<ATTACK>(em, ks, sq, bit, an & below);  // (-1,+1)
                emit_targets_asc<ATTACK>(em, ks, sq, bit, an & above);   // (+1,-1)
                emit_targets_asc<ATTACK>(em, ks, sq, bit, dg & above);   // (+1,+1)
            }
        }
    }

    if (ATTACK) return;
    // castles, lib.rs:578-610 + 966-1056: needs the mover's king on the board and K-right OR
    // Q-right (Q4); queen side is listed before king side.  The black branch tests WHITE ids
    // (+ROOK on a8/h8, +KING on e8) exactly like lib.rs:1023-1046 (Q3) -- never true in play.
    if (!ks.has_king) return;
    const u64 wR = rooks & b.w, wK = kings & b.w;
    if (white_to_move) {
        if (!(rights & (RT_WK | RT_WQ))) return;
        const u64 e1 = 1ULL << 60;
        if ((wR >> 56 & 1) && !(occ & (7ULL << 57)) && (wK & e1) && !(eatt & (7ULL << 58))) em.push(ACT_CASTLE_QS_WHITE);
        if ((wR >> 63 & 1) && !(occ & (3ULL << 61)) && (wK & e1) && !(eatt & (7ULL << 60))) em.push(ACT_CASTLE_KS_WHITE);
    } else {
        if (!(rights & (RT_BK | RT_BQ))) return;
        const u64 e8 = 1ULL << 4;
        if ((wR & 1) && !(occ & (7ULL << 1)) && (wK & e8) && !(eatt & (7ULL << 2))) em.push(ACT_CASTLE_QS_BLACK);
        if ((wR >> 7 & 1) && !(occ & (3ULL << 5)) && (wK & e8) && !(eatt & (7ULL << 4))) em.push(ACT_CASTLE_KS_BLACK);
    }
}

// ---------------------------------------------------------------------------------------------
// next_state (lib.rs:679-784) on planes.  `rights` must already be masked (mask_rights).
// Returns the reward; *status = 0 ok, -1 empty from-square (reference panics), -2 bad action.
// *irreversible is set for pawn moves and captures (board-only repetition key can never recur).
// ---------------------------------------------------------------------------------------------
GCB_HD int piece_value(int code) {
    // P1 N3 B3 R5 Q10 K0 (lib.rs:19-25), indexed by code 0..7, 4 bits each
    return (int)((0x01335A00u >> (code * 4)) & 15u);
}

GCB_HD int apply_action(Board& b, u32& rights, int white_to_move, int action, int* status,
                                            bool* irreversible) {
    *status = 0;
    *irreversible = false;
    int reward = 0;
    if (action < 4096) {
        if (action < 0) { *status = -2; return 0; }
        const int from = action >> 6, to = action & 63;
        const int pid = piece_id(b, from);
        if (pid == 0) { *status = -1; return 0; }
        const int cap = piece_code(b, to);
        const int code = pid < 0 ? -pid : pid;
        int newcode = code, newwhite = pid > 0;
        reward += piece_value(cap);
        // "pawn becomes queen" tests the WRONG ends (lib.rs:703-704, Q1); colour = the MOVER's
        if (code == PC_PAWN && ((white_to_move && (to >> 3) == 7) || (!white_to_move && (to >> 3) == 0))) {
            newcode = PC_QUEEN, newwhite = white_to_move;
            reward += 10;
        }
        clear_sq(b, 1ULL << from);
        put_sq(b, to, newcode, newwhite);
        // rights react only to WHITE ids, column of the from-square only (lib.rs:711-734, Q5)
        if (pid == PC_KING) rights &= white_to_move ? ~(RT_WK | RT_WQ) : ~(RT_BK | RT_BQ);
        else if (pid == PC_ROOK) {
            if ((from & 7) == 0) rights &= white_to_move ? ~RT_WQ : ~RT_BQ;
            else if ((from & 7) == 7) rights &= white_to_move ? ~RT_WK : ~RT_BK;
        }
        *irreversible = (code == PC_PAWN) || (cap != 0);
    } else {
        switch (action) {  // literal square writes, lib.rs:739-774
        case ACT_CASTLE_KS_WHITE:
            put_sq(b, 60, 0, 0), put_sq(b, 61, PC_ROOK, 1), put_sq(b, 62, PC_KING, 1), put_sq(b, 63, 0, 0);
            rights &= ~(RT_WK | RT_WQ);
            break;
        case ACT_CASTLE_QS_WHITE:
            put_sq(b, 56, 0, 0), put_sq(b, 57, 0, 0), put_sq(b, 58, PC_KING, 1), put_sq(b, 59, PC_ROOK, 1), put_sq(b, 60, 0, 0);
            rights &= ~(RT_WK | RT_WQ);
            break;
        case ACT_CASTLE_KS_BLACK:
            put_sq(b, 4, 0, 0), put_sq(b, 5, PC_ROOK, 0), put_sq(b, 6, PC_KING, 0), put_sq(b, 7, 0, 0);
            rights &= ~(RT_BK | RT_BQ);
            break;
        case ACT_CASTLE_QS_BLACK:
            put_sq(b, 0, 0, 0), put_sq(b, 1, 0, 0), put_sq(b, 2, PC_KING, 0), put_sq(b, 3, PC_ROOK, 0), put_sq(b, 4, 0, 0);
            rights &= ~(RT_BK | RT_BQ);
            break;
        default: *status = -2; return 0;
        }
    }
    return reward;
}

// ---------------------------------------------------------------------------------------------
// Zobrist key of the board only (the reference's repetition key is the 64-char board string,
// chess_v2.py:404-407, 599-602: no side to move, no rights).  Keys are splitmix64 of
// (piece index, square); piece index = id + 6 in 0..12.
// ---------------------------------------------------------------------------------------------
GCB_HD u64 gcb_splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
GCB_HD u64 zobrist_piece(int pid /* signed id != 0 */, int sq) {
    return gcb_splitmix64((u64)((pid + 6) * 64 + sq) + 0x6A09E667F3BCC908ULL);
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
