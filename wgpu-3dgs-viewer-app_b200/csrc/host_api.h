// host_api.h — internal declarations shared by the CUDA side (csrc/) and the host side (host/).
#pragma once
#include <stdint.h>

#include "../../include/b200gs.h"

void gs_set_error(const char* fmt, ...);
uint32_t gs_record_bytes(uint32_t sh, uint32_t cov3d);
