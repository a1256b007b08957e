// stand-in: IpoptApplication is only used by the reference main, which is not compiled here
#pragma once
#include "IpTNLP.hpp"
