// forwards to the facade (see ../../g2o_facade.hpp): the path slam.hpp:26-35 includes
#include "g2o_facade.hpp"
