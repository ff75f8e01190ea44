// Stand-in for Samples/HistogramUtils.h (ROOT histogram helpers): nothing of it is used by
// SampleHandlerBase::GetTestStatLLH / GetPoissonLLH, the functions oracle/ref_host drives.
#pragma once
