// Stand-in for Manager/MaCh3Modes.h (interaction-mode table read from YAML): SampleHandlerBase only owns a pointer.
#pragma once
class MaCh3Modes {};
