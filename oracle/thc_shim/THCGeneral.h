/* ORACLE -- TEST INFRASTRUCTURE ONLY.  See THC.h in this directory. */
#pragma once
#include "THC.h"
