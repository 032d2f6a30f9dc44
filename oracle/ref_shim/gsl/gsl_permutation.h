/* stand-in: everything the reference path needs is in ref_shim.h */
#include "../ref_shim.h"
