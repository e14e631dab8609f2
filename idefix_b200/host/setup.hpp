// setup.hpp -- the Setup class a problem's setup.cpp implements (reference: src/setup.hpp:23-28).
// The user file defines the constructor and InitFlow; the destructor has a weak default in main.cpp
// (the reference provides weak defaults the same way, src/setup.cpp:14-23).
#ifndef IDEFIX_B200_HOST_SETUP_HPP_
#define IDEFIX_B200_HOST_SETUP_HPP_
#include "idefix.hpp"

class Setup {
 public:
  Setup(Input &, Grid &, DataBlock &, Output &);
  ~Setup();
  void InitFlow(DataBlock &);
};

#endif  // IDEFIX_B200_HOST_SETUP_HPP_
