#include "inst.cuh"
namespace lq {
lq_engine* make_pend_f64(const lq_config* c, std::string& err, int* rc) { return make_for_lqr<double, PendulumSys<double>>(c, err, rc); }
}
