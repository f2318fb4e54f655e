#include "inst.cuh"
namespace lq {
lq_engine* make_boat_f64(const lq_config* c, std::string& err, int* rc) { return make_for_lqr_pd<double, BoatSys<double>>(c, err, rc); }
}
