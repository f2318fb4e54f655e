#include "inst.cuh"
namespace lq {
lq_engine* make_boat_f32(const lq_config* c, std::string& err, int* rc) { return make_for_lqr_pd<float, BoatSys<float>>(c, err, rc); }
}
