// stand-in for NuOscillator's Constants/OscillatorConstants.h (external project, absent): the flavour enum only.
#pragma once
namespace NuOscillator { enum { kElectron = 1, kMuon = 2, kTau = 3 }; }
#ifndef FLOAT_T
#define FLOAT_T double   /* NuOscillator's default build (UseDoubles) */
#endif
