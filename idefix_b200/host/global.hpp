// global.hpp -- kept so that user files written for the reference (#include "global.hpp") compile unchanged.
#include "idefix.hpp"
