// Temporary: entry points not implemented yet return an error (removed as each stage lands).
#include "internal.h"
using namespace paig;
#define NOT_YET(name) set_error(name " is not implemented yet"); return 99
extern "C" {
int paig_step_forward(const paig_task*, const paig_params*, const float*, int, const paig_outputs*, void*, void*) { NOT_YET("paig_step_forward"); }
int paig_step_backward(const paig_task*, const paig_params*, const paig_params*, const float*, int, const float*, const float*, const float*, const float*, void*, void*) { NOT_YET("paig_step_backward"); }
int paig_step_fused(const paig_task*, const paig_params*, const paig_params*, const float*, int, const paig_outputs*, void*, void*) { NOT_YET("paig_step_fused"); }
int paig_step_fused_host(const paig_task*, const paig_params*, const paig_params*, const float*, int, float*, void*, void*) { NOT_YET("paig_step_fused_host"); }
int paig_encoder_forward(const paig_task*, const paig_params*, const float*, long, int, int, float*, float*, float*, void*, void*) { NOT_YET("paig_encoder_forward"); }
int paig_encoder_backward(const paig_task*, const paig_params*, const paig_params*, const float*, long, int, int, const float*, void*, void*) { NOT_YET("paig_encoder_backward"); }
int paig_velocity_forward(const paig_task*, const paig_params*, const float*, int, float*, void*, void*) { NOT_YET("paig_velocity_forward"); }
int paig_velocity_backward(const paig_task*, const paig_params*, const paig_params*, const float*, int, const float*, float*, void*, void*) { NOT_YET("paig_velocity_backward"); }
}
