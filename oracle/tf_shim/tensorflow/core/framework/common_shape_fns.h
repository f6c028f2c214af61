// TEST INFRASTRUCTURE ONLY: forwards to the minimal TF stand-in (see tf_min.h).
#include "../../../tf_min.h"
