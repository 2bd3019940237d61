#pragma once
#include "H5File.hpp"
