// TEST INFRASTRUCTURE ONLY: see Controller.h in this directory.
#include "Controller.h"
