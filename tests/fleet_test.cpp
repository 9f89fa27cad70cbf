// Multi-GPU through the C ABI ONLY (include/grample_b200.h, no torch, no Python): one process drives n devices
// through a gb_fleet (single-process NCCL communicator) the way a Go / C++ host in the place of cmd/root.go:381-561
// would, and the results are compared with ONE device holding all the chains.  Chains are keyed by global id and the
// merge travels as 64-bit integer counts, so the merged marginals must be bit-identical; the convergence scores
// agree to rounding (their float64 sums are accumulated in a different order).
//   usage: fleet_test <dir with the .uai fixtures> [n_devices = all]
#include <cmath>
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
#define OK(call)                                                                              \
    do {                                                                                      \
        if ((call) != 0) {                                                                    \
            std::printf("  ERROR %s:%d: %s -> %s\n", __FILE__, __LINE__, #call, gb_last_error()); \
            std::exit(2);                                                                     \
        }                                                                                     \
    } while (0)

struct Run {
    std::vector<double> merged, conv;
    std::vector<int32_t> col, chosen;
    int64_t chains = 0, samples = 0;
};

// `sampler`: 0 = simple over one model, 1 = adaptive (Adapt after every round)
static Run run(const std::string& uai, const char* evid, int n_dev, const int* devices, int32_t total_chains, int precision,
               bool adaptive) {
    const uint64_t seed = 77;
    const int32_t cw = 20, rounds = 3;
    std::vector<gb_model*> models(n_dev);
    std::vector<gb_chains*> chains(n_dev);
    gb_fleet* fleet = nullptr;
    OK(gb_fleet_create(n_dev, devices, &fleet));
    const int64_t blocks = (total_chains + 7) / 8, per = (blocks + n_dev - 1) / n_dev;
    for (int i = 0; i < n_dev; i++) {
        OK(gb_model_load_uai(uai.c_str(), evid, devices[i], &models[i]));
        const int64_t first = std::min<int64_t>(i * per, blocks) * 8, last = std::min<int64_t>((i + 1) * per, blocks) * 8;
        const int32_t n = (int32_t)(std::min<int64_t>(last, total_chains) - first);
        OK(gb_chains_create(1, &models[i], &n, seed, (uint64_t)first, precision, GB_CHAINS_HISTORY, devices[i], &chains[i]));
        OK(gb_fleet_attach(fleet, i, chains[i]));
    }
    int32_t n_vars = 0, total_card = 0;
    OK(gb_model_n_vars(models[0], &n_vars));
    OK(gb_model_total_card(models[0], &total_card));
    Run r;
    r.merged.resize(total_card);
    r.col.resize(n_vars);
    r.conv.resize(n_vars);
    OK(gb_fleet_sweep(fleet, 50, 0));  // burn-in
    uint64_t next_id = (uint64_t)((total_chains + 7) / 8 * 8);
    std::vector<double> prev(total_card);
    for (int round = 0; round < rounds; round++) {
        OK(gb_fleet_advance(fleet, cw));
        // the asynchronous form: the merge of this round is read back while nothing else is pending ...
        OK(gb_fleet_merge_begin(fleet, r.merged.data(), r.col.data()));
        OK(gb_fleet_merge_end(fleet, &r.chains, &r.samples));
        // ... and must equal the blocking form
        OK(gb_fleet_merged_marginals(fleet, prev.data(), nullptr));
        EXPECT(std::memcmp(prev.data(), r.merged.data(), sizeof(double) * total_card) == 0);
        OK(gb_fleet_convergence(fleet, GB_HELLINGER, r.merged.data(), r.conv.data()));
        if (adaptive) {
            int32_t chosen[4], n_chosen = 0;
            OK(gb_fleet_adapt(fleet, models.data(), 2, 64, GB_HELLINGER, cw, 128, next_id, chosen, &n_chosen));
            for (int k = 0; k < n_chosen; k++) r.chosen.push_back(chosen[k]);
            next_id += 64ull * (uint64_t)n_chosen;
        }
    }
    OK(gb_fleet_synchronize(fleet));
    for (int i = 0; i < n_dev; i++) gb_chains_destroy(chains[i]);
    gb_fleet_destroy(fleet);
    for (int i = 0; i < n_dev; i++) gb_model_destroy(models[i]);
    return r;
}

static void compare(const char* name, const Run& one, const Run& many) {
    EXPECT(one.chains == many.chains);
    EXPECT(one.samples == many.samples);
    EXPECT(one.merged.size() == many.merged.size());
    EXPECT(std::memcmp(one.merged.data(), many.merged.data(), sizeof(double) * one.merged.size()) == 0);
    EXPECT(one.col == many.col);
    EXPECT(one.chosen == many.chosen);
    double worst = 0;
    for (size_t v = 0; v < one.conv.size(); v++) worst = std::fmax(worst, std::fabs(one.conv[v] - many.conv[v]) / std::fabs(one.conv[v]));
    EXPECT(worst < 1e-10);
    std::printf("%s %s: chains %lld, samples %lld, variants added %zu, worst relative score difference %.2e\n",
                failures ? "FAIL" : "PASS", name, (long long)many.chains, (long long)many.samples, many.chosen.size(), worst);
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::printf("usage: fleet_test <fixture dir> [n_devices]\n");
        return 2;
    }
    const std::string res = argv[1];
    int n = 0;
    OK(gb_device_count(&n));
    if (n < 1) {
        std::printf("no CUDA device: grample_b200 has no CPU fallback\n");
        return 3;
    }
    int n_dev = argc > 2 ? std::atoi(argv[2]) : n;
    if (n_dev > n) n_dev = n;
    std::vector<int> devs(n_dev);
    for (int i = 0; i < n_dev; i++) devs[i] = i;
    const int one = 0;
    std::printf("fleet_test: %d device(s)\n", n_dev);
    {
        const std::string uai = res + "/Grids_11.uai";
        const int32_t ragged = 16 * n_dev - 3;  // two blocks of 8 chains per device, the last shard ragged
        compare("TestFleetSimpleF64", run(uai, nullptr, 1, &one, ragged, GB_F64, false), run(uai, nullptr, n_dev, devs.data(), ragged, GB_F64, false));
        compare("TestFleetSimpleTable", run(uai, nullptr, 1, &one, 4096, GB_TABLE, false), run(uai, nullptr, n_dev, devs.data(), 4096, GB_TABLE, false));
    }
    {
        const std::string uai = res + "/Pedigree_11.uai", ev = uai + ".evid";
        compare("TestFleetAdaptiveHybrid", run(uai, ev.c_str(), 1, &one, 64, GB_HYBRID, true),
                run(uai, ev.c_str(), n_dev, devs.data(), 64, GB_HYBRID, true));
    }
    std::printf(failures ? "FAILED (%d)\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}
