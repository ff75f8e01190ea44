// stand-in for ROOT's Math/DistFunc.h: see root_fwd.h.
#pragma once
#include "root_fwd.h"
