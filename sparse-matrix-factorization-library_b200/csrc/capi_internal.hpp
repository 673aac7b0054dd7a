// Definitions of the opaque C-ABI objects, shared by the capi_*.cpp/.cu translation units.
#pragma once
#include <string>

#include "symbolic.hpp"

struct sfb200_symbolic {
    sfb200::Symbolic S;
};

namespace sfb200 {
void set_error(const std::string &msg);
}
