// Internal glue shared by the .cu files: the public C ABI plus the launch counter.
#pragma once
#include "../../include/season_nerf_b200.h"
namespace snb {
void count_launch(int n = 1);
}
