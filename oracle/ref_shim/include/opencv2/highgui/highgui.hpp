// TEST INFRASTRUCTURE: IngvioParams.h includes this header; none of its names is used by the compiled units.
#pragma once
