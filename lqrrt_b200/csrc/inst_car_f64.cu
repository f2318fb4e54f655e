#include "inst.cuh"
namespace lq {
lq_engine* make_car_f64(const lq_config* c, std::string& err, int* rc) { return make_for_lqr_pd<double, CarSys<double>>(c, err, rc); }
}
