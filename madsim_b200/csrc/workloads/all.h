#pragma once
#include "chaos.h"
#include "pingpong.h"
#include "raft.h"
#include "sleep.h"
#include "stream.h"
#include "etcd.h"
#include "tonic.h"
#include "signal.h"
#define MSIM_HAVE_RAFT 1
