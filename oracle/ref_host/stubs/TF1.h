// stand-in for ROOT's TF1.h (ROOT is not installed in this image): compile-only declarations, see root_fwd.h.
#pragma once
#include "root_fwd.h"
