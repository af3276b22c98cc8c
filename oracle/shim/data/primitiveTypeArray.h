/* oracle/shim forwarding header (test infrastructure): the reference includes <data/primitiveTypeArray.h>; everything lives in gegelati.h */
#include "../gegelati.h"
