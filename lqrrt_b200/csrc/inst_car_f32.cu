#include "inst.cuh"
namespace lq {
lq_engine* make_car_f32(const lq_config* c, std::string& err, int* rc) { return make_for_lqr_pd<float, CarSys<float>>(c, err, rc); }
}
