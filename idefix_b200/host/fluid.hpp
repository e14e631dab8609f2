// fluid.hpp -- kept so that user files written for the reference (#include "fluid.hpp") compile unchanged.
#include "idefix.hpp"
