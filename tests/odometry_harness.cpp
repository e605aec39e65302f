// odometry_harness.cpp -- TEST-ONLY CPU build of the product's pipeline host logic (locus_b200/csrc/odometry.cu)
// against fake stages (tests/odometry_stub.h).  Built by tests/test_odometry_cpu.py into tests/_build/.
#define LB_ODOMETRY_STUB 1
#include "odometry_stub.h"
#include "../locus_b200/csrc/odometry.cu"

extern "C" {
int oh_aligns_peak(void) { return stub::aligns_peak().load(); }
int oh_clouds_alive(void) { return clouds_alive().load(); }
int oh_clouds_peak(void) { return clouds_peak().load(); }
void oh_reset(int max_sleep_us) { clouds_peak().store(0); stub::aligns_peak().store(0); stub::max_sleep_us().store(max_sleep_us); }
}
