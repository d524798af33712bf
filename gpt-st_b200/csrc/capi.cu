// Library identification and ABI glue that does not belong to one kernel family.
#include "../../include/gptst_b200.h"

extern "C" const char* gptst_version(void) { return "gptst_b200 0.1.0 sm_100a"; }
