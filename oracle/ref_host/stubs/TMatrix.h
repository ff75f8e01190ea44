// stand-in for ROOT's TMatrix.h: see root_fwd.h.
#pragma once
#include "root_fwd.h"
