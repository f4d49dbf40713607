// Host-side plan of the intra kernel (K3): which records it owns, in which order, grouped into 64x64 luma units with the
// neighbour units each unit reads.  Pure host code (no CUDA): built by the engine while staging a frame, and checkable on its own
// (k3_plan_check, exposed as av1r_debug_k3_check for the CPU tests) because the kernel's freedom from deadlock rests on it.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "av1_consts.h"
#include "worklist.h"

namespace av1r {

inline bool k3_owns(const TxRec& r) { return r.mode != TXM_INTER || (r.flags & TXF_II); }

// 64x64 luma units [ux0, ux1] x [uy0, uy1] the predictor of an intra-block-copy record reads (same arithmetic on the device)
inline void k3_ibc_source_units(const TxRec& r, int subx, int suby, int mi_cols, int mi_rows, int& ux0, int& uy0, int& ux1, int& uy1) {
    const int sx = r.plane ? subx : 0, sy = r.plane ? suby : 0;
    const int dvx = (int16_t)r.cfl_max_w4, dvy = (int16_t)r.cfl_max_h4;       // 1/8 luma sample = 1/16 chroma sample at 4:2:0
    const int x = r.x4 * 4, y = r.y4 * 4;
    const int w = std::max(1, std::min((int)kTxW[r.txsz], ((mi_cols * 4) >> sx) - x)), h = std::max(1, std::min((int)kTxH[r.txsz], ((mi_rows * 4) >> sy) - y));
    const int posx = (x << 4) + ((2 * dvx) >> sx), posy = (y << 4) + ((2 * dvy) >> sy);   // 1/16 sample, plane units
    const int px0 = posx >> 4, py0 = posy >> 4;
    const int us = r.plane ? 6 - sx : 6, vs = r.plane ? 6 - sy : 6;
    ux0 = px0 >> us;
    uy0 = py0 >> vs;
    ux1 = (px0 + w - 1 + ((posx & 15) != 0)) >> us;     // a fractional position reads one more sample (second bilinear tap)
    uy1 = (py0 + h - 1 + ((posy & 15) != 0)) >> vs;
}

// K3 order: records grouped by 64x64 luma unit, units in wavefront order.  Unit key = 4 * (sbx + 2 * sby) + seq with (sbx, sby)
// the superblock and seq the unit's rank in decode order inside its superblock (0 for 64x64 superblocks; a 128x128 superblock
// visits its four units in Z order, or 0,2,1,3 under a vertical split).  Every sample a record may read lies earlier in its own
// unit or in a unit with a smaller key: left superblock = base - 4 (this covers the below-left samples taken from the left
// neighbour's bottom half), above-right = base - 4, above = base - 8.  The kernel hands units to CTAs in table order and waits
// only for lower table indices, so it cannot deadlock.
//   tx[n_recs]   : the frame's records in decode order
//   recs[n_recs] : the staged copy (inter-intra residual records get the K3 position of their blend record in pal_off)
//   k3[n_k3]     : out, record indices in K3 order;  units[]: out (capacity: one entry per 64x64 unit of the frame)
// Returns the number of units, or -1 if a unit holds more records than the kernel's barrier array.
// Frames that allow intra block copy (`raster` != 0) list their units in decode order instead: a block vector may point at any
// unit decoded earlier (spec 6.10.25 allows sources up to five unit columns ahead per superblock row above, which the 2:1
// wavefront order has not visited yet), and decode order is a topological order of those reads as well as of the neighbour
// reads (prediction never crosses a tile).  `upos_out` (one entry per 64x64 unit of the frame, may be null) receives the table
// index of every unit: the kernel finds the source units of a block vector through it.  Returns -2 if a block-copy record reads
// a unit that is not listed before its own.
inline int k3_plan_build(const TxRec* tx, int n_recs, int n_k3, int subx, int suby, int sb128, int mi_cols, int mi_rows, TxRec* recs,
                         uint32_t* k3, K3Unit* units, int raster = 0, int32_t* upos_out = nullptr) {
    const int sx1 = subx, sy1 = suby;
    const int sbs = sb128 ? 1 : 0;
    const int UX = (mi_cols + 15) >> 4, UY = (mi_rows + 15) >> 4;
    auto unit_of = [&](const TxRec& r) -> int {
        const int sx = r.plane ? sx1 : 0, sy = r.plane ? sy1 : 0;
        const int ux = ((r.x4 * 4) << sx) >> 6, uy = ((r.y4 * 4) << sy) >> 6;
        return std::min(uy, UY - 1) * UX + std::min(ux, UX - 1);
    };
    std::vector<int32_t> ukey((size_t)UX * UY, -1), upos((size_t)UX * UY, -1);
    std::vector<uint32_t> cnt((size_t)std::max(4096, UX * UY) + 1, 0);
    int last_sb = -1, seq = 0, rank = 0;
    for (int i = 0; i < n_recs; i++) {   // decode order: first appearance of a unit fixes its rank inside the superblock
        const TxRec& r = tx[i];
        if (!k3_owns(r)) continue;
        const int un = unit_of(r);
        if (ukey[un] < 0) {
            const int ux = un % UX, uy = un / UX;
            const int sb = (uy >> sbs) * UX + (ux >> sbs);
            seq = sb == last_sb ? std::min(seq + 1, 3) : 0;
            last_sb = sb;
            ukey[un] = raster ? rank++ : std::min(4095, 4 * ((ux >> sbs) + 2 * (uy >> sbs)) + seq);
        }
        cnt[ukey[un]]++;
    }
    uint32_t acc = 0;
    for (auto& c : cnt) { const uint32_t t = c; c = acc; acc += t; }
    for (int i = 0; i < n_recs; i++) {
        const TxRec& r = tx[i];
        if (k3_owns(r)) k3[cnt[ukey[unit_of(r)]]++] = (uint32_t)i;
    }
    // unit table: runs of equal unit in K3 order
    int nu = 0;
    for (int n = 0; n < n_k3; n++) {
        const int un = unit_of(recs[k3[n]]);
        if (nu == 0 || upos[un] != nu - 1) {
            K3Unit& u = units[nu];
            u.first = (uint32_t)n;
            u.count = 0;
            u.ux = (uint16_t)(un % UX);
            u.uy = (uint16_t)(un / UX);
            upos[un] = nu++;
        }
        units[nu - 1].count++;
    }
    // which neighbour units a unit really reads: a record on the unit's top row with an available row above reads the unit
    // above (and above-left / above-right when it touches those corners), one on the left column reads the unit to the left
    // (and below-left when the block reaches the unit's bottom).  In intra frames every neighbour is needed; in inter frames
    // the few units that hold intra / inter-intra blocks would otherwise chain up for no reason.
    std::vector<uint8_t> need((size_t)nu, 0);
    for (int k = 0; k < nu; k++) {
        const K3Unit& u = units[k];
        uint8_t nd = 0;
        for (uint32_t n = u.first; n < u.first + u.count; n++) {
            const TxRec& r = recs[k3[n]];
            if (r.mode == TXM_INTER || r.mode == TXM_PALETTE || r.mode == TXM_INTRABC) continue;
            const int sh = r.plane ? 3 : 4;                       // unit size in 4-sample cells: 16 luma, 8 chroma (4:2:0)
            const int lx = r.x4 - (u.ux << sh), ly = r.y4 - (u.uy << sh);
            const int w4 = kTxW[r.txsz] >> 2, h4 = kTxH[r.txsz] >> 2, uw = 1 << sh;
            const bool top = ly == 0 && (r.flags & TXF_HAVE_ABOVE), lft = lx == 0 && (r.flags & TXF_HAVE_LEFT);
            if (lft) nd |= 1;
            if (lft && (r.flags & TXF_HAVE_BELOW_LEFT) && ly + 2 * h4 > uw) nd |= 2;
            if ((top && lx == 0) || (lft && ly == 0)) nd |= 4;
            if (top) nd |= 8;
            if (top && (r.flags & TXF_HAVE_ABOVE_RIGHT) && lx + 2 * w4 > uw) nd |= 16;
        }
        need[k] = nd;
    }
    for (int k = 0; k < nu; k++) {
        K3Unit& u = units[k];
        static const int dxy[5][2] = {{-1, 0}, {-1, 1}, {-1, -1}, {0, -1}, {1, -1}};   // left, below-left, above-left, above, above-right
        for (int d = 0; d < 5; d++) {
            const int nx = u.ux + dxy[d][0], ny = u.uy + dxy[d][1];
            int dep = -1;
            if (((need[k] >> d) & 1) && nx >= 0 && ny >= 0 && nx < UX && ny < UY) {
                const int pos = upos[(size_t)ny * UX + nx];
                if (pos >= 0 && pos < k) dep = pos;
            }
            u.dep[d] = dep;
        }
    }
    int n_units = nu;
    for (int k = 0; k < nu; k++)
        if (units[k].count > (uint32_t)K3_UNIT_MAX_RECS) n_units = -1;   // cannot happen at 4:2:0 (<= 576 records per unit)
    // intra block copy: every unit the source rectangle touches (one sample of margin for the chroma half-sample taps) must be
    // listed before the record's own unit
    for (int k = 0; k < nu && n_units >= 0; k++) {
        const K3Unit& u = units[k];
        for (uint32_t n = u.first; n < u.first + u.count; n++) {
            const TxRec& r = recs[k3[n]];
            if (r.mode != TXM_INTRABC) continue;
            int ux0, uy0, ux1, uy1;
            k3_ibc_source_units(r, subx, suby, mi_cols, mi_rows, ux0, uy0, ux1, uy1);
            for (int yy = uy0; yy <= uy1; yy++)
                for (int xx = ux0; xx <= ux1; xx++) {
                    if (xx < 0 || yy < 0 || xx >= UX || yy >= UY) continue;   // (outside the frame: the fetch clamps)
                    const int pos = upos[(size_t)yy * UX + xx];
                    if (pos < 0 || pos >= k) n_units = -2;
                }
        }
    }
    if (upos_out) memcpy(upos_out, upos.data(), sizeof(int32_t) * (size_t)UX * UY);
    // explicit dependency of inter-intra residual records on their blend record (position in K3 order)
    uint32_t blend_pos[3] = {0, 0, 0};
    for (int n = 0; n < n_k3; n++) {
        TxRec& r = recs[k3[n]];
        if (!(r.flags & TXF_II)) continue;
        if (r.mode != TXM_INTER) blend_pos[r.plane] = (uint32_t)n;
        else r.pal_off = blend_pos[r.plane];
    }
    return n_units;
}

// Invariants the kernel relies on.  Returns an empty string or the first violation.
inline std::string k3_plan_check(const TxRec* tx, int n_recs, int n_k3, int subx, int suby, int mi_cols, int mi_rows, const TxRec* recs,
                                 const uint32_t* k3, const K3Unit* units, int n_units) {
    const int UX = (mi_cols + 15) >> 4, UY = (mi_rows + 15) >> 4;
    auto fail = [](const char* what, int a, int b) { return std::string(what) + " (" + std::to_string(a) + ", " + std::to_string(b) + ")"; };
    if (n_units < 0) return "a unit exceeds K3_UNIT_MAX_RECS";
    std::vector<int> upos((size_t)UX * UY, -1);
    uint32_t next = 0;
    std::vector<uint8_t> seen((size_t)n_recs, 0);
    for (int u = 0; u < n_units; u++) {
        const K3Unit& U = units[u];
        if (U.first != next || U.count == 0) return fail("unit ranges do not tile the K3 order", u, (int)U.first);
        next += U.count;
        if (U.ux >= UX || U.uy >= UY || upos[(size_t)U.uy * UX + U.ux] >= 0) return fail("unit position invalid or listed twice", U.ux, U.uy);
        upos[(size_t)U.uy * UX + U.ux] = u;
        uint32_t prev = 0;
        for (uint32_t n = U.first; n < U.first + U.count; n++) {
            const uint32_t idx = k3[n];
            if ((int)idx >= n_recs || seen[idx]) return fail("record listed twice or out of range", u, (int)idx);
            seen[idx] = 1;
            const TxRec& r = recs[idx];
            if (!k3_owns(r)) return fail("record not owned by K3", u, (int)idx);
            if (n > U.first && idx <= prev) return fail("records of a unit are not in decode order", u, (int)idx);
            prev = idx;
            const int sx = r.plane ? subx : 0, sy = r.plane ? suby : 0;
            if ((((r.x4 * 4) << sx) >> 6) != U.ux || (((r.y4 * 4) << sy) >> 6) != U.uy) return fail("record outside its unit", u, (int)idx);
            if (r.mode == TXM_INTRABC) {   // every source unit of a block vector comes earlier in the table
                int ux0, uy0, ux1, uy1;
                k3_ibc_source_units(r, subx, suby, mi_cols, mi_rows, ux0, uy0, ux1, uy1);
                for (int yy = uy0; yy <= uy1; yy++)
                    for (int xx = ux0; xx <= ux1; xx++) {
                        if (xx < 0 || yy < 0 || xx >= UX || yy >= UY) continue;
                        const int pos = upos[(size_t)yy * UX + xx];
                        if (pos < 0 || pos >= u) return fail("block copy reads a unit that is not finished before its own", u, (int)idx);
                    }
            }
            if ((r.flags & TXF_II) && r.mode == TXM_INTER) {   // residual of an inter-intra block: its blend record is earlier in the same unit
                if (r.pal_off < U.first || r.pal_off >= n) return fail("inter-intra residual does not follow its blend record", u, (int)idx);
                const TxRec& b = recs[k3[r.pal_off]];
                if (!(b.flags & TXF_II) || b.mode == TXM_INTER || b.plane != r.plane) return fail("inter-intra blend link broken", u, (int)idx);
            }
        }
    }
    if ((int)next != n_k3) return fail("K3 order length", (int)next, n_k3);
    for (int i = 0; i < n_recs; i++)
        if (k3_owns(tx[i]) != (seen[i] != 0)) return fail("K3 membership", i, seen[i]);
    static const int dxy[5][2] = {{-1, 0}, {-1, 1}, {-1, -1}, {0, -1}, {1, -1}};
    for (int u = 0; u < n_units; u++) {
        const K3Unit& U = units[u];
        for (int d = 0; d < 5; d++) {
            const int nx = U.ux + dxy[d][0], ny = U.uy + dxy[d][1];
            const int pos = (nx >= 0 && ny >= 0 && nx < UX && ny < UY) ? upos[(size_t)ny * UX + nx] : -1;
            if (U.dep[d] >= 0 && (U.dep[d] != pos || pos >= u)) return fail("dependency is not the lower-index neighbour", u, d);
        }
        // every neighbour a record of the unit can read (edge-availability flags) that is listed earlier must be a dependency
        for (uint32_t n = U.first; n < U.first + U.count; n++) {
            const TxRec& r = recs[k3[n]];
            if (r.mode == TXM_INTER || r.mode == TXM_PALETTE || r.mode == TXM_INTRABC) continue;
            const int sh = r.plane ? 3 : 4, uw = 1 << sh;
            const int lx = r.x4 - (U.ux << sh), ly = r.y4 - (U.uy << sh);
            const int w4 = kTxW[r.txsz] >> 2, h4 = kTxH[r.txsz] >> 2;
            bool need[5] = {false, false, false, false, false};
            const bool top = ly == 0 && (r.flags & TXF_HAVE_ABOVE), lft = lx == 0 && (r.flags & TXF_HAVE_LEFT);
            need[0] = lft;
            need[1] = lft && (r.flags & TXF_HAVE_BELOW_LEFT) && ly + 2 * h4 > uw;
            need[2] = top && lft;
            need[3] = top;
            need[4] = top && (r.flags & TXF_HAVE_ABOVE_RIGHT) && lx + 2 * w4 > uw;
            for (int d = 0; d < 5; d++) {
                if (!need[d]) continue;
                const int nx = U.ux + dxy[d][0], ny = U.uy + dxy[d][1];
                if (nx < 0 || ny < 0 || nx >= UX || ny >= UY) continue;
                const int pos = upos[(size_t)ny * UX + nx];
                if (pos >= 0 && pos < u && U.dep[d] != pos) return fail("a neighbour unit that is read is not a dependency", u, d);
                if (pos > u) return fail("a record reads a unit that comes later in the order", u, d);
            }
        }
    }
    return std::string();
}

}  // namespace av1r
