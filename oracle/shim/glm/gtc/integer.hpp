// Empty on purpose: the reference includes this header (stdafx.h:22) but uses nothing from it on the hot path.
#pragma once
