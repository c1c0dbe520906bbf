// session.cu -- device-resident streaming session (placeholder until the fused path lands).
#include "../../include/ryk.h"
#include "engine.h"
#include "features.h"

namespace ryk {
void session_destroy_all(Engine* e) { (void)e; }
}

extern "C" {
int ryk_session_create(ryk_engine*, const ryk_session_config*, int*) { ryk::set_error("sessions not implemented yet"); return -3; }
int ryk_session_destroy(ryk_engine*, int) { ryk::set_error("sessions not implemented yet"); return -3; }
int ryk_session_push(ryk_engine*, int, const float*, int, double*, int, int*) { ryk::set_error("sessions not implemented yet"); return -3; }
int ryk_session_push_device(ryk_engine*, int, const float*, int, double*, int, int*) { ryk::set_error("sessions not implemented yet"); return -3; }
}
