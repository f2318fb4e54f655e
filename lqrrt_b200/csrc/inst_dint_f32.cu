#include "inst.cuh"
namespace lq {
lq_engine* make_dint_f32(const lq_config* c, std::string& err, int* rc) { return make_dint<float>(c, err, rc); }
}
