// sparse_lu.h -- symbolic phase of the fixed-pattern sparse factorisation of the Newton matrix
// I - gamma*J of the large-model path (host code, run once per model at pb200_create).
//
// What it replaces in the reference: SUNLinSol_KLU's one-time analysis
// (deps/AMICI/ThirdParty/sundials/src/sunlinsol/klu/sunlinsol_klu.c:259-330: klu_analyze --
// BTF + fill-reducing ordering -- on the first setup; every later setup is klu_refactor on the
// fixed pattern and pivot order; every Newton iteration is klu_solve), AMICI's default linear
// solver (deps/AMICI/include/amici/solver.h, LinearSolver::KLU).  Here:
//
//   * ordering: minimum degree on the pattern of J + J^T (ties by index, deterministic);
//   * split: the first n1 pivots -- those whose degree in the elimination graph stays <= deg_max --
//     form the SPARSE part, eliminated with their diagonal as pivot (KLU's own preference: it keeps
//     the diagonal unless |diag| < 0.001 * max|column|, and refactor never re-pivots); what remains
//     is the m x m Schur complement, which for networks without small separators fills in almost
//     completely and is therefore handled DENSE by the tensor-core path of bdf_team.cuh
//     (partial pivoting inside the block);
//   * the numeric phase on the device is an entry-wise (left-looking) evaluation
//         LU(i,j) = M(i,j) - sum_{k < min(i,j)} L(i,k) U(k,j)        [L entries then / U(j,j)]
//     with a precomputed (L position, U position) pair list per target entry, the targets ordered
//     by dependency LEVEL: every entry is written once by one thread, products are summed in
//     ascending pivot order -- no atomics, run-to-run reproducible;
//   * triangular solves are row-wise gathers over dependency levels of the rows, all right-hand
//     sides (1 + nplist columns) side by side.
//
// All row/column numbers in the tables are PERMUTED indices (position in the elimination order)
// unless stated otherwise.  Values live in one array V per condition:
//     V = [ 1/U(k,k), k < n1 | L entries, row by row | U entries (strict upper), row by row ].
#ifndef PB200_SPARSE_LU_H
#define PB200_SPARSE_LU_H

#include <stdint.h>

#include <algorithm>
#include <cstdint>
#include <vector>

namespace pb200 {

/// Device (or host-harness) view of the symbolic factorisation: plain pointers.
struct SparseLU {
    int enabled;   // 0: no symbolic factorisation (dense path only)
    int nx, n1, m; // sparse pivots, dense Schur block (n1 + m = nx)
    int nV, nnzL, nnzU;
    const int* ord;   // [nx] permuted -> original index
    const int* iord;  // [nx] original -> permuted index
    // numeric phase: targets (V entries) sorted by level
    int n_tlevels;
    const int* tlevel_ptr;  // [n_tlevels + 1]
    const int* t_dst;       // V index written
    const int* t_src;       // ELL position of the Jacobian entry it starts from, -1: structural fill
    const int* t_piv;       // -1: U entry; -2: diagonal; >= 0: L entry, V index of 1/pivot of its column
    const int* t_pptr;      // [n_targets + 1] pair lists
    const int* p_l;         // V index of L(i,k)
    const int* p_u;         // V index of U(k,j)
    // Schur block targets (entries of S that are not structurally zero before the dense phase)
    int n_starg;
    const int* s_dst;   // position in S: row + col * m
    const int* s_src;   // ELL position or -1
    const int* s_diag;  // 1: diagonal entry of the matrix (identity contributes 1)
    const int* s_pptr;  // [n_starg + 1], pairs continue in p_l / p_u
    // row-wise views for the triangular solves
    // (16-bit column / row numbers: nx < 65536; the team kernel keeps these tables in shared memory)
    const int* Lr_ptr;  // [nx + 1] L entries of permuted row r (columns k < min(r, n1)); value at V[n1 + e]
    const unsigned short* Lr_col;
    const int* Ur_ptr;  // [n1 + 1] strict-upper entries of row r < n1; value at V[n1 + nnzL + e]
    const unsigned short* Ur_col;
    int n_flev, n_blev;
    const int* flev_ptr;   // forward levels over the sparse rows
    const unsigned short* flev_rows;
    const int* blev_ptr;   // backward levels
    const unsigned short* blev_rows;
    const unsigned short* iord16;  // iord again, 16 bit (staging loops: address and source select without a global load)
    // per level: 1 = its rows are long (mean >= 48 entries): a row's chain is split over 16 lanes instead of
    // being walked by one thread ([n_flev] / [n_blev]; entry n_flev of flev_long: the Schur rows)
    const unsigned char* flev_long;
    const unsigned char* blev_long;
    const unsigned char* tlev_long;  // [n_tlevels] numeric phase: the level's targets have long pair lists (mean >= 64)
    /// bytes of the solve tables (row pointers + 16-bit indices), the part worth keeping in shared memory
    int solve_table_bytes() const {
        return 4 * ((nx + 1) + (n1 + 1) + (n_flev + 1) + (n_blev + 1)) + 2 * (nnzL + nnzU + 2 * n1 + nx);
    }
};

struct SparseLUHost {
    int nx = 0, n1 = 0, m = 0, nV = 0, nnzL = 0, nnzU = 0, nnzJ = 0;
    std::vector<int> ord, iord;
    std::vector<int> tlevel_ptr, t_dst, t_src, t_piv, t_pptr, p_l, p_u;
    std::vector<int> s_dst, s_src, s_diag, s_pptr;
    std::vector<int> Lr_ptr, Lr_col, Ur_ptr, Ur_col;
    std::vector<int> flev_ptr, flev_rows, blev_ptr, blev_rows;
    std::vector<unsigned short> Lr_col16, Ur_col16, flev_rows16, blev_rows16, iord16;  // what the solves read
    std::vector<unsigned char> flev_long, blev_long, tlev_long;
    int64_t n_pairs_sparse = 0, n_pairs_schur = 0;

    /// view over the host vectors (host harness)
    SparseLU view() const {
        SparseLU s{};
        s.enabled = 1;
        s.nx = nx; s.n1 = n1; s.m = m; s.nV = nV; s.nnzL = nnzL; s.nnzU = nnzU;
        s.ord = ord.data(); s.iord = iord.data();
        s.n_tlevels = (int)tlevel_ptr.size() - 1;
        s.tlevel_ptr = tlevel_ptr.data(); s.t_dst = t_dst.data(); s.t_src = t_src.data(); s.t_piv = t_piv.data();
        s.t_pptr = t_pptr.data(); s.p_l = p_l.data(); s.p_u = p_u.data();
        s.n_starg = (int)s_dst.size();
        s.s_dst = s_dst.data(); s.s_src = s_src.data(); s.s_diag = s_diag.data(); s.s_pptr = s_pptr.data();
        s.Lr_ptr = Lr_ptr.data(); s.Lr_col = Lr_col16.data(); s.Ur_ptr = Ur_ptr.data(); s.Ur_col = Ur_col16.data();
        s.n_flev = (int)flev_ptr.size() - 1; s.n_blev = (int)blev_ptr.size() - 1;
        s.flev_ptr = flev_ptr.data(); s.flev_rows = flev_rows16.data();
        s.blev_ptr = blev_ptr.data(); s.blev_rows = blev_rows16.data();
        s.iord16 = iord16.data();
        s.flev_long = flev_long.data(); s.blev_long = blev_long.data(); s.tlev_long = tlev_long.data();
        return s;
    }

    /// flops of one numeric factorisation of the sparse part + assembly of the Schur block
    /// (the dense phase on the m x m block is counted by the caller)
    int64_t factor_flops() const { return 2 * (n_pairs_sparse + n_pairs_schur) + nnzL + n1; }
    /// flops of the two sparse triangular solves for ONE right-hand side
    int64_t solve_flops() const { return 2 * ((int64_t)nnzL + nnzU) + n1; }

    /// Symbolic analysis of the pattern (CSR j_ptr / j_col of the nx x nx Jacobian, ell_pos[e] = position
    /// of CSR entry e in the ELL value array).  deg_max: largest elimination-graph degree of a sparse
    /// pivot; m_align: the dense block's dimension is rounded up to a multiple of it.
    static SparseLUHost build(int nx, const int* j_ptr, const int* j_col, const int* ell_pos, int deg_max,
                              int m_align = 2, int m_target = 0) {
        SparseLUHost H;
        H.nx = nx;
        H.nnzJ = j_ptr[nx];
        const int W = (nx + 63) / 64;
        auto bit_set = [&](std::vector<uint64_t>& b, int row, int c) { b[(size_t)row * W + (c >> 6)] |= (uint64_t)1 << (c & 63); };
        auto bit_get = [&](const std::vector<uint64_t>& b, int row, int c) {
            return (b[(size_t)row * W + (c >> 6)] >> (c & 63)) & 1;
        };
        // ---- 1. minimum degree on J + J^T
        std::vector<uint64_t> adj((size_t)nx * W, 0);
        for (int i = 0; i < nx; ++i)
            for (int e = j_ptr[i]; e < j_ptr[i + 1]; ++e) {
                const int j = j_col[e];
                if (j != i) { bit_set(adj, i, j); bit_set(adj, j, i); }
            }
        std::vector<char> alive(nx, 1);
        std::vector<int> deg(nx);
        auto popc = [&](int row) {
            int s = 0;
            for (int w = 0; w < W; ++w) s += __builtin_popcountll(adj[(size_t)row * W + w]);
            return s;
        };
        for (int i = 0; i < nx; ++i) deg[i] = popc(i);
        H.ord.clear();
        int n_alive = nx;
        int extra = 0, over_budget = 0;
        while (n_alive > 0) {
            int best = -1;
            for (int i = 0; i < nx; ++i)
                if (alive[i] && (best < 0 || deg[i] < deg[best])) best = i;
            // past deg_max the elimination goes on only to bring the dense block down to a multiple of
            // m_align (full warps / tensor-core tiles of the dense path), for at most m_align - 1 pivots of
            // at most twice that degree
            if (m_target > 0) {  // (eliminate down to exactly m_target rows, whatever the degrees)
                if (n_alive <= m_target) break;
            } else if (deg[best] > deg_max) {
                const int over = n_alive % (m_align > 1 ? m_align : 1);
                if (extra == 0) over_budget = over;
                if (over == 0 || deg[best] > 2 * deg_max || extra >= over_budget) break;
                ++extra;
            }
            // eliminate `best`: its neighbours become a clique
            const int k = best;
            alive[k] = 0;
            --n_alive;
            H.ord.push_back(k);
            std::vector<int> nb;
            for (int j = 0; j < nx; ++j)
                if (bit_get(adj, k, j)) nb.push_back(j);
            for (int a : nb) {
                uint64_t* ra = &adj[(size_t)a * W];
                const uint64_t* rk = &adj[(size_t)k * W];
                for (int w = 0; w < W; ++w) ra[w] |= rk[w];
                ra[a >> 6] &= ~((uint64_t)1 << (a & 63));
                ra[k >> 6] &= ~((uint64_t)1 << (k & 63));
                deg[a] = popc(a);
            }
        }
        // the dense block: what is left, in ascending original index; its dimension a multiple of m_align
        int n1 = (int)H.ord.size();
        if (extra > 0 && m_align > 1 && (nx - n1) % m_align != 0) n1 -= extra;  // alignment not reached: undo the extras
        if (m_align > 1) {
            int m = nx - n1;
            const int m_up = ((m + m_align - 1) / m_align) * m_align;
            int give_back = m_up - m;
            if (give_back > n1) give_back = n1;
            n1 -= give_back;  // the LAST sparse pivots (highest degrees) go back to the dense block
        }
        std::vector<char> in_sparse(nx, 0);
        H.ord.resize(n1);
        for (int k : H.ord) in_sparse[k] = 1;
        for (int i = 0; i < nx; ++i)
            if (!in_sparse[i]) H.ord.push_back(i);
        H.n1 = n1;
        H.m = nx - n1;
        H.iord.assign(nx, 0);
        for (int r = 0; r < nx; ++r) H.iord[H.ord[r]] = r;

        // ---- 2. exact symbolic elimination of the permuted (unsymmetric) pattern
        std::vector<uint64_t> R((size_t)nx * W, 0), Cc((size_t)nx * W, 0);  // row / column patterns
        std::vector<int> src_of;  // ELL position of original entries, looked up through a per-row map below
        for (int i = 0; i < nx; ++i) {
            const int r = H.iord[i];
            bit_set(R, r, r); bit_set(Cc, r, r);  // the identity: diagonal always present
            for (int e = j_ptr[i]; e < j_ptr[i + 1]; ++e) {
                const int c = H.iord[j_col[e]];
                bit_set(R, r, c); bit_set(Cc, c, r);
            }
        }
        std::vector<int> Lk, Uk;
        for (int k = 0; k < n1; ++k) {
            Lk.clear(); Uk.clear();
            for (int j = k + 1; j < nx; ++j) {
                if (bit_get(R, k, j)) Uk.push_back(j);
                if (bit_get(Cc, k, j)) Lk.push_back(j);
            }
            for (int i : Lk)
                for (int j : Uk) { bit_set(R, i, j); bit_set(Cc, j, i); }
        }
        // ---- 3. row-wise L and U
        H.Lr_ptr.assign(nx + 1, 0);
        H.Ur_ptr.assign(n1 + 1, 0);
        for (int r = 0; r < nx; ++r) {
            const int lim = r < n1 ? r : n1;
            for (int k = 0; k < lim; ++k)
                if (bit_get(R, r, k)) H.Lr_col.push_back(k);
            H.Lr_ptr[r + 1] = (int)H.Lr_col.size();
        }
        for (int r = 0; r < n1; ++r) {
            for (int j = r + 1; j < nx; ++j)
                if (bit_get(R, r, j)) H.Ur_col.push_back(j);
            H.Ur_ptr[r + 1] = (int)H.Ur_col.size();
        }
        H.nnzL = (int)H.Lr_col.size();
        H.nnzU = (int)H.Ur_col.size();
        H.nV = n1 + H.nnzL + H.nnzU;
        const int m = H.m;
        // V index of entry (r, c) with min(r, c) < n1, or -1
        auto vpos = [&](int r, int c) -> int {
            if (r == c) return r < n1 ? r : -1;
            if (c < r) {  // L
                const int* b = H.Lr_col.data() + H.Lr_ptr[r];
                const int* e = H.Lr_col.data() + H.Lr_ptr[r + 1];
                const int* it = std::lower_bound(b, e, c);
                return (it != e && *it == c) ? n1 + (int)(it - H.Lr_col.data()) : -1;
            }
            if (r >= n1) return -1;
            const int* b = H.Ur_col.data() + H.Ur_ptr[r];
            const int* e = H.Ur_col.data() + H.Ur_ptr[r + 1];
            const int* it = std::lower_bound(b, e, c);
            return (it != e && *it == c) ? n1 + H.nnzL + (int)(it - H.Ur_col.data()) : -1;
        };
        // original entries: ELL source position per permuted (r, c)
        std::vector<int> vsrc(H.nV, -1);
        std::vector<int> ssrc((size_t)m * m, -1);
        std::vector<char> stouched((size_t)m * m, 0);
        for (int i = 0; i < nx; ++i)
            for (int e = j_ptr[i]; e < j_ptr[i + 1]; ++e) {
                const int r = H.iord[i], c = H.iord[j_col[e]];
                if (r >= n1 && c >= n1) {
                    ssrc[(size_t)(r - n1) + (size_t)(c - n1) * m] = ell_pos[e];
                    stouched[(size_t)(r - n1) + (size_t)(c - n1) * m] = 1;
                } else {
                    vsrc[vpos(r, c)] = ell_pos[e];
                }
            }
        for (int s = 0; s < m; ++s) stouched[(size_t)s + (size_t)s * m] = 1;  // identity
        // ---- 4. pair lists per target, ascending pivot
        std::vector<std::vector<std::pair<int, int>>> vpairs(H.nV);
        std::vector<std::vector<std::pair<int, int>>> spairs;  // created lazily through smap
        std::vector<int> smap((size_t)m * m, -1);
        for (int k = 0; k < n1; ++k) {
            // column k of L: rows i > k with (i, k) in L; row k of U
            Lk.clear();
            for (int i = k + 1; i < nx; ++i)
                if (bit_get(R, i, k)) Lk.push_back(i);
            for (int i : Lk) {
                const int lp = vpos(i, k);
                for (int e = H.Ur_ptr[k]; e < H.Ur_ptr[k + 1]; ++e) {
                    const int j = H.Ur_col[e];
                    const int up = n1 + H.nnzL + e;
                    if (i >= n1 && j >= n1) {
                        const size_t sp = (size_t)(i - n1) + (size_t)(j - n1) * m;
                        if (smap[sp] < 0) { smap[sp] = (int)spairs.size(); spairs.emplace_back(); }
                        spairs[smap[sp]].emplace_back(lp, up);
                        stouched[sp] = 1;
                        ++H.n_pairs_schur;
                    } else {
                        vpairs[vpos(i, j)].emplace_back(lp, up);
                        ++H.n_pairs_sparse;
                    }
                }
            }
        }
        // ---- 5. dependency levels of the V targets
        std::vector<int> lev(H.nV, 0);
        // process in an order in which sources precede targets: by min(r, c) ascending, U entries and
        // the diagonal of a pivot before the L entries of its column
        struct Ent { int r, c, v; };
        std::vector<Ent> ents;
        for (int r = 0; r < nx; ++r) {
            if (r < n1) ents.push_back({r, r, r});
            for (int e = H.Lr_ptr[r]; e < H.Lr_ptr[r + 1]; ++e) ents.push_back({r, H.Lr_col[e], n1 + e});
            if (r < n1)
                for (int e = H.Ur_ptr[r]; e < H.Ur_ptr[r + 1]; ++e) ents.push_back({r, H.Ur_col[e], n1 + H.nnzL + e});
        }
        std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) {
            const int ka = a.r < a.c ? a.r : a.c, kb = b.r < b.c ? b.r : b.c;
            if (ka != kb) return ka < kb;
            const int ca = a.r > a.c, cb = b.r > b.c;  // L entries last within a pivot
            if (ca != cb) return ca < cb;
            if (a.r != b.r) return a.r < b.r;
            return a.c < b.c;
        });
        int maxlev = 0;
        for (const Ent& en : ents) {
            int l = 0;
            for (auto& pr : vpairs[en.v]) l = std::max(l, std::max(lev[pr.first], lev[pr.second]) + 1);
            if (en.r > en.c) l = std::max(l, lev[en.c] + 1);  // L entry: needs the pivot of its column
            lev[en.v] = l;
            maxlev = std::max(maxlev, l);
        }
        std::vector<int> order(ents.size());
        for (size_t i = 0; i < ents.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lev[ents[a].v] < lev[ents[b].v]; });
        H.tlevel_ptr.assign(maxlev + 2, 0);
        H.t_pptr.push_back(0);
        for (int oi : order) {
            const Ent& en = ents[oi];
            H.t_dst.push_back(en.v);
            H.t_src.push_back(vsrc[en.v]);
            H.t_piv.push_back(en.r == en.c ? -2 : (en.r > en.c ? en.c : -1));
            for (auto& pr : vpairs[en.v]) { H.p_l.push_back(pr.first); H.p_u.push_back(pr.second); }
            H.t_pptr.push_back((int)H.p_l.size());
            H.tlevel_ptr[lev[en.v] + 1]++;
        }
        for (int l = 0; l <= maxlev; ++l) H.tlevel_ptr[l + 1] += H.tlevel_ptr[l];
        if (ents.empty()) H.tlevel_ptr.assign(1, 0);
        // ---- 6. Schur targets
        H.s_pptr.push_back((int)H.p_l.size());
        for (int c = 0; c < m; ++c)
            for (int r = 0; r < m; ++r) {
                const size_t sp = (size_t)r + (size_t)c * m;
                if (!stouched[sp]) continue;
                H.s_dst.push_back((int)sp);
                H.s_src.push_back(ssrc[sp]);
                H.s_diag.push_back(r == c ? 1 : 0);
                if (smap[sp] >= 0)
                    for (auto& pr : spairs[smap[sp]]) { H.p_l.push_back(pr.first); H.p_u.push_back(pr.second); }
                H.s_pptr.push_back((int)H.p_l.size());
            }
        // ---- 7. levels of the triangular solves
        {
            std::vector<int> fl(n1, 0);
            int mx = -1;
            for (int r = 0; r < n1; ++r) {
                int l = 0;
                for (int e = H.Lr_ptr[r]; e < H.Lr_ptr[r + 1]; ++e) l = std::max(l, fl[H.Lr_col[e]] + 1);
                fl[r] = l;
                mx = std::max(mx, l);
            }
            H.flev_ptr.assign(mx + 2, 0);
            for (int r = 0; r < n1; ++r) H.flev_ptr[fl[r] + 1]++;
            for (int l = 0; l <= mx; ++l) H.flev_ptr[l + 1] += H.flev_ptr[l];
            H.flev_rows.assign(n1, 0);
            std::vector<int> fill(H.flev_ptr.begin(), H.flev_ptr.end());
            for (int r = 0; r < n1; ++r) H.flev_rows[fill[fl[r]]++] = r;
            std::vector<int> bl(n1, 0);
            mx = -1;
            for (int r = n1 - 1; r >= 0; --r) {
                int l = 0;
                for (int e = H.Ur_ptr[r]; e < H.Ur_ptr[r + 1]; ++e)
                    if (H.Ur_col[e] < n1) l = std::max(l, bl[H.Ur_col[e]] + 1);
                bl[r] = l;
                mx = std::max(mx, l);
            }
            H.blev_ptr.assign(mx + 2, 0);
            for (int r = 0; r < n1; ++r) H.blev_ptr[bl[r] + 1]++;
            for (int l = 0; l <= mx; ++l) H.blev_ptr[l + 1] += H.blev_ptr[l];
            H.blev_rows.assign(n1, 0);
            fill.assign(H.blev_ptr.begin(), H.blev_ptr.end());
            for (int r = 0; r < n1; ++r) H.blev_rows[fill[bl[r]]++] = r;
            if (n1 == 0) { H.flev_ptr.assign(1, 0); H.blev_ptr.assign(1, 0); }
        }
        H.Lr_col16.assign(H.Lr_col.begin(), H.Lr_col.end());
        H.Ur_col16.assign(H.Ur_col.begin(), H.Ur_col.end());
        H.flev_rows16.assign(H.flev_rows.begin(), H.flev_rows.end());
        H.blev_rows16.assign(H.blev_rows.begin(), H.blev_rows.end());
        H.iord16.assign(H.iord.begin(), H.iord.end());
        {
            const int nfl = (int)H.flev_ptr.size() - 1, nbl = (int)H.blev_ptr.size() - 1;
            H.flev_long.assign(nfl + 1, 0);
            H.blev_long.assign(nbl + 1, 0);
            for (int l = 0; l < nfl; ++l) {
                int64_t tot = 0;
                const int nr = H.flev_ptr[l + 1] - H.flev_ptr[l];
                for (int e = H.flev_ptr[l]; e < H.flev_ptr[l + 1]; ++e) tot += H.Lr_ptr[H.flev_rows[e] + 1] - H.Lr_ptr[H.flev_rows[e]];
                H.flev_long[l] = (nr > 0 && tot >= 48 * (int64_t)nr) ? 1 : 0;
            }
            if (m > 0) H.flev_long[nfl] = (H.Lr_ptr[nx] - H.Lr_ptr[n1] >= 48 * (int64_t)m) ? 1 : 0;
            const int ntl = (int)H.tlevel_ptr.size() - 1;
            H.tlev_long.assign(ntl + 1, 0);
            for (int l = 0; l < ntl; ++l) {
                const int t0 = H.tlevel_ptr[l], t1 = H.tlevel_ptr[l + 1];
                const int64_t tot = (int64_t)H.t_pptr[t1] - H.t_pptr[t0];
                H.tlev_long[l] = (t1 > t0 && tot >= 64 * (int64_t)(t1 - t0)) ? 1 : 0;
            }
            for (int l = 0; l < nbl; ++l) {
                int64_t tot = 0;
                const int nr = H.blev_ptr[l + 1] - H.blev_ptr[l];
                for (int e = H.blev_ptr[l]; e < H.blev_ptr[l + 1]; ++e) tot += H.Ur_ptr[H.blev_rows[e] + 1] - H.Ur_ptr[H.blev_rows[e]];
                H.blev_long[l] = (nr > 0 && tot >= 48 * (int64_t)nr) ? 1 : 0;
            }
        }
        return H;
    }

    /// The split the team kernel uses: sparse pivots up to degree 32; if the dense block that leaves is
    /// larger than `rows_max` (the dense path's fast variant owns one matrix row per thread), the sparse
    /// elimination is carried on in minimum-degree order until exactly rows_max rows are left -- as long as
    /// the numeric phase stays below `max_pairs` multiply-adds.  Small models (nx < 128): degree limit nx / 4.
    static SparseLUHost build_for_team(int nx, const int* j_ptr, const int* j_col, const int* ell_pos, int rows_max = 512,
                                       int64_t max_pairs = 6000000) {
        if (nx < 128) return build(nx, j_ptr, j_col, ell_pos, nx / 4 > 2 ? nx / 4 : 2, 2);
        SparseLUHost best = build(nx, j_ptr, j_col, ell_pos, 32, 32);
        if (best.m > rows_max) {
            SparseLUHost next = build(nx, j_ptr, j_col, ell_pos, nx, 32, rows_max);
            if (next.n_pairs_sparse + next.n_pairs_schur <= max_pairs) best = std::move(next);
        }
        return best;
    }
};

}  // namespace pb200
#endif
