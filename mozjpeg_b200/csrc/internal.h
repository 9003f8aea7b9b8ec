// internal.h -- shared declarations inside libb200jpeg (not installed).
#pragma once
#include "b200jpeg.h"
namespace b200 {
void set_error(const char *fmt, ...);
}
