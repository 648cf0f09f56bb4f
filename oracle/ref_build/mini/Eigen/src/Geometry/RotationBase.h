// part of Core in this stand-in (oracle test infrastructure)
#pragma once
#include "../../Core"
