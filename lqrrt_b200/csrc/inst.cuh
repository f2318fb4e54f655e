// One translation unit per (vehicle, dtype) so the model instantiations build in parallel.
#pragma once
#include "engine.cuh"

namespace lq {
lq_engine* make_boat_f32(const lq_config*, std::string&, int*);
lq_engine* make_boat_f64(const lq_config*, std::string&, int*);
lq_engine* make_car_f32(const lq_config*, std::string&, int*);
lq_engine* make_car_f64(const lq_config*, std::string&, int*);
lq_engine* make_pend_f32(const lq_config*, std::string&, int*);
lq_engine* make_pend_f64(const lq_config*, std::string&, int*);
lq_engine* make_dint_f32(const lq_config*, std::string&, int*);
lq_engine* make_dint_f64(const lq_config*, std::string&, int*);

template <typename T, class Sys>
lq_engine* make_for_lqr(const lq_config* c, std::string& err, int* rc) {
    switch (c->lqr_kind) {
    case LQ_LQR_CONSTANT: return make_engine<T, Sys, ConstantLqr<T, Sys>>(c, err, rc);
    case LQ_LQR_FD_RICCATI: return make_engine<T, Sys, FdRiccatiLqr<T, Sys>>(c, err, rc);
    default: err = "this dynamics model does not support the requested lqr provider"; *rc = LQ_ERR_UNSUPPORTED; return nullptr;
    }
}
template <typename T, class Sys>
lq_engine* make_for_lqr_pd(const lq_config* c, std::string& err, int* rc) {
    if (c->lqr_kind == LQ_LQR_PD_ROTATED) return make_engine<T, Sys, PdRotatedLqr<T, Sys>>(c, err, rc);
    return make_for_lqr<T, Sys>(c, err, rc);
}
template <typename T>
lq_engine* make_dint(const lq_config* c, std::string& err, int* rc) {
    switch (c->n) {
    case 4: return make_for_lqr<T, DoubleIntSys<T, 2>>(c, err, rc);
    case 6: return make_for_lqr<T, DoubleIntSys<T, 3>>(c, err, rc);
    case 12: return make_for_lqr<T, DoubleIntSys<T, 6>>(c, err, rc);
    default: err = "double integrator is compiled for 2, 3 and 6 axes"; *rc = LQ_ERR_UNSUPPORTED; return nullptr;
    }
}
}  // namespace lq
