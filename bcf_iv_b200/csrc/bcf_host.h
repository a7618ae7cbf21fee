// bcf_host.h -- host-side passes compiled by the host compiler alone (bcf_host.cpp)
#pragma once
#include <cstdint>

namespace bcf {
// min, max of a column and flags: bit 0 = every value equals the min or the max (and all are finite), bit 1 = NaN / Inf present
void scan_column(const double* x, int64_t n, double* lo_out, double* hi_out, uint8_t* flag_out);
}
