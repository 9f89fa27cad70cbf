// Thread safety of ONE chain handle ("every entry thread-safe per handle"): the Go shim's goroutines call
// gb_chains_group_advance / gb_chains_synchronize / gb_chains_group_counts / gb_chains_get_state on one handle
// concurrently (go/sampler_cuda.go; the reference spawns one goroutine per chain, chain.go:197-215).  16 pthreads
// do exactly that here, each owning one group; the result must equal the same calls made from a single thread
// (groups are independent, so any interleaving of whole calls gives the same trajectories and counts).
//   usage: threads_test <dir with the .uai fixtures>
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/grample_b200.h"

static int failures = 0;
#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) {                                                        \
            std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
            failures++;                                                       \
        }                                                                     \
    } while (0)
#define OK(call)                                                                                  \
    do {                                                                                          \
        if ((call) != 0) {                                                                        \
            std::printf("  ERROR %s:%d: %s -> %s\n", __FILE__, __LINE__, #call, gb_last_error()); \
            std::exit(2);                                                                         \
        }                                                                                         \
    } while (0)

constexpr int kThreads = 16, kChains = 24, kRounds = 6, kCw = 10;

struct Shared {
    gb_chains* chains;
    int32_t n_vars, total_card;
    std::vector<std::vector<uint64_t>> counts;  // per group
    std::vector<std::vector<int32_t>> state;    // per group
    std::vector<int64_t> samples;
};
struct Arg {
    Shared* sh;
    int group;
};

static void drive_group(Shared* sh, int g) {
    for (int r = 0; r < kRounds; r++) {
        OK(gb_chains_group_advance(sh->chains, g, kCw));
        OK(gb_chains_synchronize(sh->chains));
        OK(gb_chains_group_counts(sh->chains, g, sh->counts[g].data()));
        OK(gb_chains_get_state(sh->chains, g, sh->state[g].data()));
        int32_t n = 0;
        OK(gb_chains_group_info(sh->chains, g, &n, &sh->samples[g], nullptr));
    }
}
static void* worker(void* p) {
    Arg* a = static_cast<Arg*>(p);
    drive_group(a->sh, a->group);
    return nullptr;
}

static Shared make(gb_model* m, int precision) {
    Shared sh{};
    std::vector<gb_model*> models(kThreads, m);
    std::vector<int32_t> per(kThreads, kChains);
    OK(gb_chains_create(kThreads, models.data(), per.data(), 5, 0, precision, GB_CHAINS_HISTORY, 0, &sh.chains));
    OK(gb_model_n_vars(m, &sh.n_vars));
    OK(gb_model_total_card(m, &sh.total_card));
    sh.counts.assign(kThreads, std::vector<uint64_t>(sh.total_card));
    sh.state.assign(kThreads, std::vector<int32_t>((size_t)kChains * sh.n_vars));
    sh.samples.assign(kThreads, 0);
    return sh;
}

static void test(const std::string& uai, int precision, const char* name) {
    const int before = failures;
    gb_model* m = nullptr;
    OK(gb_model_load_uai(uai.c_str(), nullptr, 0, &m));
    Shared serial = make(m, precision), par = make(m, precision);
    for (int g = 0; g < kThreads; g++) drive_group(&serial, g);
    pthread_t th[kThreads];
    Arg args[kThreads];
    for (int g = 0; g < kThreads; g++) {
        args[g] = Arg{&par, g};
        pthread_create(&th[g], nullptr, worker, &args[g]);
    }
    for (int g = 0; g < kThreads; g++) pthread_join(th[g], nullptr);
    for (int g = 0; g < kThreads; g++) {
        EXPECT(serial.counts[g] == par.counts[g]);
        EXPECT(serial.state[g] == par.state[g]);
        EXPECT(serial.samples[g] == par.samples[g] && par.samples[g] > 0);
    }
    std::vector<double> m1(serial.total_card), m2(serial.total_card);
    OK(gb_chains_merged_marginals(serial.chains, m1.data(), nullptr));
    OK(gb_chains_merged_marginals(par.chains, m2.data(), nullptr));
    EXPECT(m1 == m2);
    std::vector<double> c1(serial.n_vars), c2(serial.n_vars);
    OK(gb_chains_convergence(serial.chains, GB_HELLINGER, m1.data(), c1.data()));
    OK(gb_chains_convergence(par.chains, GB_HELLINGER, m2.data(), c2.data()));
    for (int v = 0; v < serial.n_vars; v++) EXPECT(std::abs(c1[v] - c2[v]) <= 1e-10 * std::abs(c1[v]));
    gb_chains_destroy(serial.chains);
    gb_chains_destroy(par.chains);
    gb_model_destroy(m);
    std::printf("%s %s\n", failures == before ? "PASS" : "FAIL", name);
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::printf("usage: threads_test <fixture dir>\n");
        return 2;
    }
    int n = 0;
    OK(gb_device_count(&n));
    if (n < 1) {
        std::printf("no CUDA device: grample_b200 has no CPU fallback\n");
        return 3;
    }
    const std::string res = argv[1];
    test(res + "/Grids_11.uai", GB_F64, "TestThreadsF64");
    test(res + "/Grids_11.uai", GB_TABLE, "TestThreadsTable");
    test(res + "/ObjectDetection_11.uai", GB_F32, "TestThreadsF32HighCard");
    std::printf(failures ? "FAILED (%d)\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}
