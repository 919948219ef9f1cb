// tools/spla_algos.cpp -- BASELINE configs 1 and 3 through spla's OWN public C++ API (spla.hpp, reference src/algorithm.cpp), once
// on spla's CPU backend (Library::set_force_no_acceleration(true), the pattern of examples/bfs.cpp:85-107) and once on the CUDA
// backend of this repository, in one process, with the results compared bit for bit and both timed on the same box:
//   sssp  FLOAT PLUS/MIN on a side x side 4-neighbour grid (road-like), symmetric weights in [1, 2), source 0
//   bfs   INT BAND/BOR/EQZERO on an R-MAT graph, push-pull (front factor 0.05), highest-degree source
// usage: spla_algos [grid_side=1024] [rmat_scale=16] [cpu=1]      (cpu=0 skips the CPU pass: the 4096-side grid takes minutes there)
// Built by spla_b200/src/Makefile into spla_b200/lib/spla_algos (needs the reference checkout at build time only).
#include <spla.hpp>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

using namespace spla;
using Clock = std::chrono::steady_clock;

static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

struct Coo {
    uint              n = 0;
    std::vector<uint> Ai, Aj;
};

static Coo make_grid(uint side) {
    Coo g;
    g.n = side * side;
    for (uint y = 0; y < side; ++y)
        for (uint x = 0; x < side; ++x) {// row-sorted, ascending columns: up, left, right, down
            const uint i = y * side + x;
            if (y > 0) g.Ai.push_back(i), g.Aj.push_back(i - side);
            if (x > 0) g.Ai.push_back(i), g.Aj.push_back(i - 1);
            if (x + 1 < side) g.Ai.push_back(i), g.Aj.push_back(i + 1);
            if (y + 1 < side) g.Ai.push_back(i), g.Aj.push_back(i + side);
        }
    return g;
}

static Coo make_rmat(int scale, int edge_factor, unsigned seed) {
    Coo                                   g;
    g.n = 1u << scale;
    std::mt19937                          rng(seed);
    std::uniform_real_distribution<float> uni(0.f, 1.f);
    std::vector<unsigned long long>       keys;
    for (std::size_t e = 0; e < std::size_t(edge_factor) << scale; ++e) {
        unsigned long long i = 0, j = 0;
        for (int b = 0; b < scale; ++b) {
            const float r = uni(rng);
            i = (i << 1) | (r >= 0.76f);
            j = (j << 1) | ((r >= 0.57f && r < 0.76f) || r >= 0.95f);
        }
        if (i == j) continue;
        keys.push_back(i * g.n + j);
        keys.push_back(j * g.n + i);
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    for (auto k : keys) g.Ai.push_back(uint(k / g.n)), g.Aj.push_back(uint(k % g.n));
    return g;
}

template<typename T>
static ref_ptr<Matrix> build_matrix(const Coo& g, const ref_ptr<Type>& type, const std::vector<T>& values) {
    auto A = Matrix::make(g.n, g.n, type);
    A->build(MemView::make((void*) g.Ai.data(), g.Ai.size() * sizeof(uint)), MemView::make((void*) g.Aj.data(), g.Aj.size() * sizeof(uint)),
             MemView::make((void*) values.data(), values.size() * sizeof(T)));
    return A;
}

template<typename T>
static std::vector<T> read_all(const ref_ptr<Vector>& v) {
    std::vector<T> out(v->get_n_rows());
    for (uint i = 0; i < out.size(); ++i) {
        if constexpr (std::is_same<T, int>::value) v->get_int(i, out[i]);
        else v->get_float(i, out[i]);
    }
    return out;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const uint side  = argc > 1 ? uint(std::atoi(argv[1])) : 1024u;
    const int  scale = argc > 2 ? std::atoi(argv[2]) : 16;
    const bool cpu   = argc > 3 ? std::atoi(argv[3]) != 0 : true;
    Library*    lib = Library::get();
    std::string info;
    lib->get_accelerator_info(info);
    std::printf("accelerator: %s\n", info.c_str());
    if (info.find("CUDA") == std::string::npos && info.find("cuda") == std::string::npos) {
        std::printf("the CUDA accelerator is not active\n");
        return 2;
    }
    int failed = 0;
    // ---- config 3: sssp on the grid --------------------------------------------------------------------------------------------
    if (side > 1) {
        Coo                g = make_grid(side);
        std::vector<float> w(g.Ai.size());
        for (std::size_t k = 0; k < w.size(); ++k) {
            const unsigned long long lo = std::min(g.Ai[k], g.Aj[k]), hi = std::max(g.Ai[k], g.Aj[k]);
            w[k] = 1.0f + float((lo * 2654435761ull + hi * 40503ull) % 1000003ull) / 1000003.0f;
        }
        auto t0 = Clock::now();
        auto A  = build_matrix<float>(g, FLOAT, w);
        std::printf("sssp grid %ux%u: n=%u nnz=%zu, Matrix::build %.1f ms\n", side, side, g.n, g.Ai.size(), ms_since(t0));
        auto desc = Descriptor::make();
        desc->set_traversal_mode(Descriptor::TraversalMode::PushPull);
        desc->set_front_factor(0.05f);
        std::vector<float> res[2];
        for (int pass = cpu ? 0 : 1; pass < 2; ++pass) {
            lib->set_force_no_acceleration(pass == 0);
            double best = 1e300, first = 0;
            for (int rep = 0; rep < (pass == 0 ? 1 : 3); ++rep) {
                auto v = Vector::make(g.n, FLOAT);
                t0     = Clock::now();
                sssp(v, A, 0, desc);
                const double ms = ms_since(t0);
                if (rep == 0) first = ms;
                best = std::min(best, ms);
                if (rep == 0) res[pass] = read_all<float>(v);
            }
            std::printf("  sssp %-4s first call %.1f ms (includes the format conversions), best %.1f ms = %.1f us per front (%u fronts)\n",
                        pass == 0 ? "cpu" : "cuda", first, best, best * 1e3 / (2.0 * (side - 1) + 1.0), 2 * (side - 1) + 1);
        }
        if (cpu) {
            const bool same = std::memcmp(res[0].data(), res[1].data(), res[0].size() * sizeof(float)) == 0;
            std::printf("  [%s] sssp grid: cuda distances == cpu distances (bit-exact)\n", same ? " OK " : "FAIL");
            failed += !same;
        }
        std::printf("  max distance %.3f\n", *std::max_element(res[1].begin(), res[1].end()));
    }
    // ---- config 1: bfs on R-MAT -----------------------------------------------------------------------------------------------
    if (scale > 0) {
        Coo               g = make_rmat(scale, 16, 1);
        std::vector<int>  ones(g.Ai.size(), 1);
        std::vector<uint> deg(g.n, 0);
        for (uint i : g.Ai) ++deg[i];
        const uint source = uint(std::max_element(deg.begin(), deg.end()) - deg.begin());
        auto       A      = build_matrix<int>(g, INT, ones);
        std::printf("bfs rmat-%d: n=%u nnz=%zu source %u\n", scale, g.n, g.Ai.size(), source);
        auto desc = Descriptor::make();
        desc->set_traversal_mode(Descriptor::TraversalMode::PushPull);
        desc->set_front_factor(0.05f);
        std::vector<int> res[2];
        for (int pass = cpu ? 0 : 1; pass < 2; ++pass) {
            lib->set_force_no_acceleration(pass == 0);
            double best = 1e300, first = 0;
            for (int rep = 0; rep < 5; ++rep) {
                auto v = Vector::make(g.n, INT);
                auto t0 = Clock::now();
                bfs(v, A, source, desc);
                const double ms = ms_since(t0);
                if (rep == 0) first = ms;
                best = std::min(best, ms);
                if (rep == 0) res[pass] = read_all<int>(v);
            }
            std::size_t edges = 0;
            for (uint i = 0; i < g.n; ++i)
                if (res[pass][i] > 0) edges += deg[i];
            std::printf("  bfs %-4s first call %.2f ms, best %.3f ms = %.3f GTEPS (entries of reached rows / time)\n", pass == 0 ? "cpu" : "cuda", first, best,
                        double(edges) / best / 1e6);
        }
        if (cpu) {
            const bool same = res[0] == res[1];
            std::printf("  [%s] bfs: cuda depths == cpu depths (bit-exact)\n", same ? " OK " : "FAIL");
            failed += !same;
        }
    }
    return failed ? 1 : 0;
}
