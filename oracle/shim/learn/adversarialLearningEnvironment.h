/* oracle/shim forwarding header (test infrastructure): the reference includes <learn/adversarialLearningEnvironment.h>; everything lives in gegelati.h */
#include "../gegelati.h"
