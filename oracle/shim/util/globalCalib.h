#pragma once
#include "util/settings.h"
