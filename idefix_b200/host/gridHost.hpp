// gridHost.hpp -- kept so that user files written for the reference (#include "gridHost.hpp") compile unchanged.
#include "idefix.hpp"
