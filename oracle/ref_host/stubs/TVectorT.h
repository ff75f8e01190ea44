// stand-in for ROOT's TVectorT.h (ROOT is not installed in this image): compile-only declarations, see root_fwd.h.
#pragma once
#include "root_fwd.h"
