// stand-in for ROOT's TLatex.h (ROOT is not installed in this image): compile-only declarations, see root_fwd.h.
#pragma once
#include "root_fwd.h"
