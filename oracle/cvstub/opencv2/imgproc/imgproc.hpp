#include "cvstub.hpp"
