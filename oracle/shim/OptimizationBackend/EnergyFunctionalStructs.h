#pragma once
#include "FullSystem/HessianBlocks.h"
#include "util/globalFuncs.h"
