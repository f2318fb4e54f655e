#include "inst.cuh"
namespace lq {
lq_engine* make_pend_f32(const lq_config* c, std::string& err, int* rc) { return make_for_lqr<float, PendulumSys<float>>(c, err, rc); }
}
