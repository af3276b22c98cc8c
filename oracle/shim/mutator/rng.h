/* oracle/shim forwarding header (test infrastructure): the reference includes <mutator/rng.h>; everything lives in gegelati.h */
#include "../gegelati.h"
