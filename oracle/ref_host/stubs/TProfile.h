// stand-in for ROOT's TProfile.h (ROOT is not installed in this image): deliberately empty.
// The reference code compiled through oracle/ref_host uses nothing from it.
#pragma once
