// ORACLE — test infrastructure only: see ../core.hpp.
#pragma once
#include "../core.hpp"
