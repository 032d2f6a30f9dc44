#include "ref_shim.h"
