// the reference includes this OptiX-SDK header from sceneLoader.h but uses nothing of it
#pragma once
