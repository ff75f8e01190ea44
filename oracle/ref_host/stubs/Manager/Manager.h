// Stand-in for Manager/Manager.h (the YAML-driven run manager: needs yaml-cpp and ROOT).  The reference sources
// compiled through oracle/ref_host only need what Manager.h pulls in transitively: logger, exception, constants.
// TEST INFRASTRUCTURE; found before the reference's own header because -Istubs precedes -I/root/reference.
#pragma once
#include <algorithm>
#include <iostream>
#include <memory>
#include "Manager/MaCh3Logger.h"
#include "Manager/MaCh3Exception.h"
#include "Manager/Core.h"
