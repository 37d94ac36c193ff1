// Opaque handle types shared by the host-side C ABI translation units.
#pragma once
#include <memory>

#include "mpp.h"

struct mppb_ptg
{
    std::unique_ptr<mpp::ptg_t> obj;
};
