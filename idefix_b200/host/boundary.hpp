// boundary.hpp -- kept so that user files written for the reference (#include "boundary.hpp") compile unchanged.
#include "idefix.hpp"
