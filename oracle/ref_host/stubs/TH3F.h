// stand-in for ROOT's TH3F.h: see root_fwd.h.
#pragma once
#include "root_fwd.h"
