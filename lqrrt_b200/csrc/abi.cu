// extern "C" surface of include/lqrrt_b200.h: argument checks, model dispatch, no exceptions across the boundary.
#include "inst.cuh"

#include <map>
#include <mutex>
#include <tuple>

static thread_local std::string g_create_err = "";
static std::mutex g_plugin_mu;
static std::map<std::tuple<int, int, int>, lq_plugin_factory>& plugin_registry() {
    static std::map<std::tuple<int, int, int>, lq_plugin_factory> r;
    return r;
}

extern "C" {

int lq_abi_version(void) { return LQ_ABI_VERSION; }

int lq_create(const lq_config* cfg, lq_engine** out) {
    if (!cfg || !out) { g_create_err = "lq_create: null argument"; return LQ_ERR_ARG; }
    *out = nullptr;
    if (cfg->abi_version != LQ_ABI_VERSION) { g_create_err = "lq_create: ABI version mismatch"; return LQ_ERR_ARG; }
    if (cfg->n < 1 || cfg->n > LQ_MAX_N || cfg->m < 1 || cfg->m > LQ_MAX_M) { g_create_err = "lq_create: n/m out of range"; return LQ_ERR_ARG; }
    if (cfg->dtype != LQ_F32 && cfg->dtype != LQ_F64) { g_create_err = "lq_create: dtype must be LQ_F32 or LQ_F64"; return LQ_ERR_ARG; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev < 1) {
        g_create_err = std::string("lq_create: no CUDA device (") + cudaGetErrorString(ce) + "); this engine has no CPU path";
        return LQ_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "lq_create: device ordinal out of range"; return LQ_ERR_ARG; }
    const bool f64 = cfg->dtype == LQ_F64;
    int rc = LQ_ERR_UNSUPPORTED;
    std::string err = "lq_create: unknown dynamics kind";
    lq_engine* e = nullptr;
    if (cfg->sys_kind >= LQ_SYS_PLUGIN_BASE) {          // user-compiled dynamics: the plugin library's own factory
        lq_plugin_factory make = nullptr;
        {
            std::lock_guard<std::mutex> g(g_plugin_mu);
            auto it = plugin_registry().find(std::make_tuple(cfg->sys_kind, cfg->dtype, cfg->lqr_kind));
            if (it != plugin_registry().end()) make = it->second;
        }
        if (!make) { g_create_err = "lq_create: no plugin registered for this (sys_kind, dtype, lqr_kind)"; return LQ_ERR_UNSUPPORTED; }
        char buf[512] = {0};
        rc = LQ_ERR_STATE;
        e = make(cfg, buf, (int)sizeof buf, &rc);
        if (!e) { g_create_err = std::string("lq_create (plugin): ") + buf; return rc == LQ_OK ? LQ_ERR_STATE : rc; }
        *out = e;
        return LQ_OK;
    }
    try {
        switch (cfg->sys_kind) {
        case LQ_SYS_BOAT: e = f64 ? lq::make_boat_f64(cfg, err, &rc) : lq::make_boat_f32(cfg, err, &rc); break;
        case LQ_SYS_CAR: e = f64 ? lq::make_car_f64(cfg, err, &rc) : lq::make_car_f32(cfg, err, &rc); break;
        case LQ_SYS_PENDULUM2: e = f64 ? lq::make_pend_f64(cfg, err, &rc) : lq::make_pend_f32(cfg, err, &rc); break;
        case LQ_SYS_DOUBLE_INTEGRATOR: e = f64 ? lq::make_dint_f64(cfg, err, &rc) : lq::make_dint_f32(cfg, err, &rc); break;
        default: break;
        }
    } catch (const std::exception& ex) {
        err = std::string("lq_create: ") + ex.what(); rc = LQ_ERR_STATE; e = nullptr;
    }
    if (!e) { g_create_err = err; return rc; }
    *out = e;
    return LQ_OK;
}

int lq_register_plugin(int sys_kind, int dtype, int lqr_kind, lq_plugin_factory make) {
    if (sys_kind < LQ_SYS_PLUGIN_BASE || !make || (dtype != LQ_F32 && dtype != LQ_F64)) return LQ_ERR_ARG;
    std::lock_guard<std::mutex> g(g_plugin_mu);
    plugin_registry()[std::make_tuple(sys_kind, dtype, lqr_kind)] = make;
    return LQ_OK;
}

int lq_destroy(lq_engine* h) { delete h; return LQ_OK; }

const char* lq_last_error(const lq_engine* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

#define LQ_H(h) do { if (!(h)) return LQ_ERR_ARG; } while (0)
#define LQ_TRY(expr) try { return (expr); } catch (const std::exception& ex) { h->err = ex.what(); return LQ_ERR_STATE; }

int lq_set_goal(lq_engine* h, const double* g, const double* b) { LQ_H(h); if (!g || !b) return LQ_ERR_ARG; LQ_TRY(h->set_goal(g, b)) }
int lq_set_horizon(lq_engine* h, int hz, void* st) { LQ_H(h); LQ_TRY(h->set_horizon(hz, (cudaStream_t)st)) }
int lq_set_sampling(lq_engine* h, const double* lo, const double* hi, const double* bias, int tries) {
    LQ_H(h); if (!lo || !hi || !bias) return LQ_ERR_ARG; LQ_TRY(h->set_sampling(lo, hi, bias, tries)) }
int lq_reset(lq_engine* h, const double* x0, void* st) { LQ_H(h); if (!x0) return LQ_ERR_ARG; LQ_TRY(h->reset(x0, (cudaStream_t)st)) }
int lq_load_nodes(lq_engine* h, int count, const double* s, const int32_t* p, const int32_t* c, void* st) {
    LQ_H(h); if (!s) return LQ_ERR_ARG; LQ_TRY(h->load_nodes(count, s, p, c, (cudaStream_t)st)) }
int lq_tree_size(lq_engine* h, int* out) { LQ_H(h); if (!out) return LQ_ERR_ARG; LQ_TRY(h->tree_size(out)) }
int lq_nearest(lq_engine* h, const void* q, int W, int ign, int32_t* idx, void* cost, void* st) {
    LQ_H(h); if (!q || !idx) return LQ_ERR_ARG; LQ_TRY(h->nearest(q, W, ign, idx, cost, (cudaStream_t)st)) }
int lq_steer(lq_engine* h, const int32_t* ids, const void* t, int W, void* x, void* u, int32_t* len, void* st) {
    LQ_H(h); if (!ids || !t) return LQ_ERR_ARG; LQ_TRY(h->steer(ids, t, W, x, u, len, (cudaStream_t)st)) }
int lq_wave_expand(lq_engine* h, uint64_t seed, uint32_t wave, int first, int count, const double* hs, int ign, int flags, void* st) {
    LQ_H(h); LQ_TRY(h->wave_expand(seed, wave, first, count, hs, ign, flags, (cudaStream_t)st)) }
int lq_wave_expand_candidates(lq_engine* h, int count, int tries, const double* c, int nc, int ign, int* used, void* st) {
    LQ_H(h); if (!c) return LQ_ERR_ARG; LQ_TRY(h->wave_expand_candidates(count, tries, c, nc, ign, used, (cudaStream_t)st)) }
int lq_wave_records(lq_engine* h, void** d, uint64_t* bytes) { LQ_H(h); LQ_TRY(h->wave_records(d, bytes)) }
int lq_wave_commit(lq_engine* h, const void* rec, int nseg, int myseg, void* st) { LQ_H(h); LQ_TRY(h->wave_commit(rec, nseg, myseg, (cudaStream_t)st)) }
int lq_wave_stats_get(lq_engine* h, lq_wave_stats* out) { LQ_H(h); LQ_TRY(h->wave_stats(out)) }
int lq_plan_run(lq_engine* h, const lq_plan_args* a, lq_plan_summary* out, void* st) { LQ_H(h); if (!a || !out) return LQ_ERR_ARG; LQ_TRY(h->plan_run(a, out, (cudaStream_t)st)) }
int lq_plan_readout(lq_engine* h, int node, const double* xg, const uint8_t* mask, int max_len, int32_t* ids, int* len,
                    double* x, double* u, int32_t* elen, void* st) {
    LQ_H(h); if (!ids || !len || (x && (!u || !elen))) return LQ_ERR_ARG;
    LQ_TRY(h->plan_readout(node, xg, mask, max_len, ids, len, x, u, elen, (cudaStream_t)st)) }
int lq_policy_wave(int wave_fixed, int wave_min, int wave_max, double growth, int tree_size) {
    lq::WavePolicy p; memset(&p, 0, sizeof p);
    p.fixed = wave_fixed; p.wmin = wave_min; p.wmax = wave_max; p.growth = growth;
    return lq::policy_wave(p, tree_size);
}
int lq_policy_slice(int W, int rank, int world, int32_t* first, int32_t* count) {
    if (!first || !count || world < 1 || rank < 0 || rank >= world || W < 0) return LQ_ERR_ARG;
    int f = 0, c = 0;
    lq::policy_slice(W, rank, world, &f, &c);
    *first = f; *count = c;
    return LQ_OK;
}
int lq_stop_word(lq_engine* h, int32_t** out) { LQ_H(h); if (!out) return LQ_ERR_ARG; LQ_TRY(h->stop_word(out)) }
int lq_closest(lq_engine* h, const double* xg, const uint8_t* mask, int* id, void* st) {
    LQ_H(h); if (!xg || !id) return LQ_ERR_ARG; LQ_TRY(h->closest(xg, mask, id, (cudaStream_t)st)) }
int lq_steer_force(lq_engine* h, int node, const double* xt, int ms, double* x, double* u, int* len, void* st) {
    LQ_H(h); if (!xt || !x || !u || !len) return LQ_ERR_ARG; LQ_TRY(h->steer_force(node, xt, ms, x, u, len, (cudaStream_t)st)) }
int lq_append_node(lq_engine* h, int parent, const double* s, const double* x, const double* u, int len, void* st) {
    LQ_H(h); if (!s || (len > 0 && (!x || !u))) return LQ_ERR_ARG; LQ_TRY(h->append_node(parent, s, x, u, len, (cudaStream_t)st)) }
int lq_mark_ignore(lq_engine* h, const int32_t* ids, int count, void* st) { LQ_H(h); LQ_TRY(h->mark_ignore(ids, count, (cudaStream_t)st)) }
int lq_climb(lq_engine* h, int node, int32_t* ids, int max_len, int* len, void* st) {
    LQ_H(h); if (!ids || !len) return LQ_ERR_ARG; LQ_TRY(h->climb(node, ids, max_len, len, (cudaStream_t)st)) }
int lq_download_nodes(lq_engine* h, int first, int count, double* s, int32_t* p, int32_t* c, int32_t* el, uint8_t* ig, uint8_t* ow, double* K, void* st) {
    LQ_H(h); LQ_TRY(h->download_nodes(first, count, s, p, c, el, ig, ow, K, (cudaStream_t)st)) }
int lq_download_edges(lq_engine* h, const int32_t* ids, int count, double* x, double* u, int32_t* len, void* st) {
    LQ_H(h); if (count > 0 && (!ids || !x || !u)) return LQ_ERR_ARG; LQ_TRY(h->download_edges(ids, count, x, u, len, (cudaStream_t)st)) }
int lq_eval_dynamics(lq_engine* h, int count, const double* x, const double* u, double dt, double* xn) {
    LQ_H(h); if (!x || !u || !xn) return LQ_ERR_ARG; LQ_TRY(h->eval_dynamics(count, x, u, dt, xn)) }
int lq_eval_lqr(lq_engine* h, int count, const double* x, const double* u, double* S, double* K) {
    LQ_H(h); if (!x || !u) return LQ_ERR_ARG; LQ_TRY(h->eval_lqr(count, x, u, S, K)) }
int lq_eval_erf(lq_engine* h, int count, const double* xg, const double* x, double* e) {
    LQ_H(h); if (!xg || !x || !e) return LQ_ERR_ARG; LQ_TRY(h->eval_erf(count, xg, x, e)) }
int lq_eval_feasible(lq_engine* h, int count, const double* x, const double* u, uint8_t* ok) {
    LQ_H(h); if (!x || !u || !ok) return LQ_ERR_ARG; LQ_TRY(h->eval_feasible(count, x, u, ok)) }
int lq_launch_count(lq_engine* h, uint64_t* out) { LQ_H(h); if (!out) return LQ_ERR_ARG; *out = h->launches; return LQ_OK; }
int lq_set_tuning(lq_engine* h, int steer_lanes, int nearest_q) { LQ_H(h); LQ_TRY(h->set_tuning(steer_lanes, nearest_q)) }
int lq_set_option(lq_engine* h, int opt, int v) { LQ_H(h); LQ_TRY(h->set_option(opt, v)) }
int lq_get_option(lq_engine* h, int opt, int* v) { LQ_H(h); if (!v) return LQ_ERR_ARG; LQ_TRY(h->get_option(opt, v)) }
int lq_wave_set_peers(lq_engine* h, int world, int rank, void* const* bufs, uint64_t bytes) { LQ_H(h); LQ_TRY(h->set_peers(world, rank, bufs, bytes)) }
int lq_set_profiling(lq_engine* h, int on) { LQ_H(h); LQ_TRY(h->set_profiling(on)) }
int lq_profile_read(lq_engine* h, double* ms, double* work, int reset) { LQ_H(h); LQ_TRY(h->profile_read(ms, work, reset)) }

}  // extern "C"
