// Version and error-string entry points of the C ABI.
#include "kernels.h"

extern "C" {
const char* rb200_version(void) { return "riporb200 0.1.0 (sm_100a)"; }
const char* rb200_last_error(void) { return rb::last_error_slot().c_str(); }
int rb200_relative_position_bucket(int relative_position, int bidirectional, int num_buckets, int max_distance) {
  return rb::relative_bucket(relative_position, bidirectional != 0, num_buckets, max_distance);
}
}
