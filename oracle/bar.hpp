// ORACLE — TEST INFRASTRUCTURE ONLY (see adiff.hpp). Bar3D / SoilContact restatement: filled in below.
#pragma once
#include "rotations.hpp"
