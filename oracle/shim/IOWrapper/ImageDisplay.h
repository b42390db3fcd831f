#pragma once
#include "IOWrapper/Output3DWrapper.h"
