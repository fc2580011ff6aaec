// Simulation of the bank-aware placement for one scan step: SC = 32*CPL chunks, rows of L chunks (8 entries each),
// random columns.  Strategies: 0 = row-supply priority (current), 1 = global-excess priority, 2 = + double groups.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
using namespace std;
int CPL = 8, SC, NG;
struct Step {
    vector<int> seg_of_chunk;          // SC
    vector<vector<int>> sup;           // [seg][32]
    int nseg;
};
static bool try_kuhn(int l, const vector<int>& ls, const Step& st, const vector<long>& prio, unsigned elig, vector<int>& bl, vector<int>& lb, unsigned& visited) {
    // banks in priority order
    int r = ls[l];
    vector<pair<long,int>> cand;
    for (int b = 0; b < 32; ++b) if (((elig >> b) & 1) && st.sup[r][b] > 0 && !((visited >> b) & 1)) cand.push_back({-(prio[b] * 4096 + st.sup[r][b]), b});
    sort(cand.begin(), cand.end());
    for (auto& c : cand) {
        int b = c.second;
        if ((visited >> b) & 1) continue;
        visited |= 1u << b;
        if (bl[b] < 0 || try_kuhn(bl[b], ls, st, prio, elig, bl, lb, visited)) { bl[b] = l; lb[l] = b; return true; }
    }
    return false;
}
int main(int argc, char** argv) {
    int strategy = argc > 1 ? atoi(argv[1]) : 0;
    CPL = argc > 2 ? atoi(argv[2]) : 8;
    int L = argc > 3 ? atoi(argv[3]) : 15;
    int trials = argc > 4 ? atoi(argv[4]) : 200;
    SC = 32 * CPL; NG = 8 * CPL;
    mt19937 rng(123);
    double tot_wave = 0, tot_bound = 0; long tot_g = 0;
    for (int t = 0; t < trials; ++t) {
        Step st; st.seg_of_chunk.resize(SC);
        int off = rng() % L, seg = 0;
        for (int c = 0; c < SC; ++c) { st.seg_of_chunk[c] = seg; if ((c + off) % L == L - 1) ++seg; }
        st.nseg = st.seg_of_chunk[SC - 1] + 1;
        st.sup.assign(st.nseg, vector<int>(32, 0));
        vector<int> gl(32, 0);
        for (int c = 0; c < SC; ++c) for (int e = 0; e < 8; ++e) { int b = rng() % 32; st.sup[st.seg_of_chunk[c]][b]++; gl[b]++; }
        tot_bound += *max_element(gl.begin(), gl.end());
        int gleft = NG;
        for (int j = 0; j < 8; ++j) for (int q = 0; q < CPL; ++q, --gleft) {
            vector<int> ls(32); for (int l = 0; l < 32; ++l) ls[l] = st.seg_of_chunk[CPL * l + q];
            vector<int> mult(32, 0), lb(32, -1);
            vector<long> prio(32, 0);
            bool dbl = false;
            if (strategy >= 1) for (int b = 0; b < 32; ++b) { prio[b] = gl[b] - gleft + 2048; if (gl[b] > gleft) dbl = true; }
            auto run_match = [&](unsigned elig, vector<int>& lanes) {
                vector<int> bl(32, -1);
                for (int l : lanes) { unsigned vis = 0; try_kuhn(l, ls, st, prio, elig, bl, lb, vis); }
            };
            vector<int> lanes(32); for (int l = 0; l < 32; ++l) lanes[l] = l;
            auto commit = [&]() { for (int l = 0; l < 32; ++l) if (lb[l] >= 0) { st.sup[ls[l]][lb[l]]--; gl[lb[l]]--; mult[lb[l]]++; ls[l] = -2 - ls[l]; } };
            if (strategy == 2 && dbl) {
                unsigned elig = 0; for (int b = 0; b < 32; ++b) if (gl[b] > gleft) elig |= 1u << b;
                run_match(elig, lanes);
                // commit matched, then second matching for the rest
                vector<int> rest; vector<int> lb1 = lb;
                for (int l = 0; l < 32; ++l) if (lb[l] >= 0) { st.sup[ls[l]][lb[l]]--; gl[lb[l]]--; mult[lb[l]]++; } else rest.push_back(l);
                for (int l = 0; l < 32; ++l) lb[l] = -1;
                for (int b = 0; b < 32; ++b) prio[b] = gl[b] - gleft + 2048;
                { vector<int> bl(32, -1); for (int l : rest) { unsigned vis = 0; try_kuhn(l, ls, st, prio, 0xffffffffu, bl, lb, vis); } }
                for (int l : rest) if (lb[l] >= 0) { st.sup[ls[l]][lb[l]]--; gl[lb[l]]--; mult[lb[l]]++; }
                for (int l = 0; l < 32; ++l) if (lb1[l] >= 0) lb[l] = lb1[l];
            } else {
                run_match(0xffffffffu, lanes);
                for (int l = 0; l < 32; ++l) if (lb[l] >= 0) { st.sup[ls[l]][lb[l]]--; gl[lb[l]]--; mult[lb[l]]++; }
            }
            for (int l = 0; l < 32; ++l) if (lb[l] < 0) {   // collide: min multiplicity, then larger excess
                int r = ls[l], best = -1;
                for (int b = 0; b < 32; ++b) if (st.sup[r][b] > 0) if (best < 0 || mult[b] < mult[best] || (mult[b] == mult[best] && gl[b] > gl[best])) best = b;
                if (best >= 0) { st.sup[r][best]--; gl[best]--; mult[best]++; lb[l] = best; }
            }
            tot_wave += *max_element(mult.begin(), mult.end()); ++tot_g;
        }
    }
    printf("strategy %d CPL %d L %d: %.3f wavefronts/gather (bound %.3f)\n", strategy, CPL, L, tot_wave / tot_g, tot_bound / (double)(trials * NG));
}
