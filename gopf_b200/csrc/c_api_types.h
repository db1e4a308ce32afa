// Handle layouts shared by the C-ABI translation units.
#pragma once
#include "model.h"

struct gopf_model {
    gopf::Model m;
    int live_solvers = 0;
};
