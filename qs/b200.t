-- qs/b200.t -- the Terra side of the B200 many-chain MCMC engine.
--
-- Drop-in for the hot path of qs/mcmc.t: `qs.b200.MCMC(kernel, params)` has the signature and the
-- return shape of `qs.MCMC(kernel, params)` (qs/mcmc.t:36-105): it returns `function(program)` which
-- returns `terra(samples: &S.Vector(infer.SampleType(program)))`.  The kernel constructors
-- `qs.b200.HMCKernel{...}`, `DriftKernel{...}`, `HARMKernel{...}`, `MixtureKernel(kernels, weights)` take the
-- parameter tables of qs/hmc.t:52-66, qs/drift.t:55-73, qs/harm.t:12-22, qs/mcmc.t:199-203 and yield a
-- POD spec the CUDA library reads, instead of a `function(TraceType) -> struct`.
--
-- The library is reached the way the reference reaches C: terralib.includec on a plain-C header
-- (qs/lib/ad.t:3, qs/lib/random.t:2-7, qs/lib/util.t:45-61).  A non-zero return becomes the reference's
-- print + abort (S.assert, qs/lib/std.t:35-43).
--
-- NOTE: this image has no Terra compiler, so this file is shipped un-run; the same ABI is exercised from
-- Python (quicksand_b200/_lib.py) and C (tests/test_abi.py::test_header_is_plain_c).  INTEGRATION.md shows
-- the two-line change to qs/init.t that registers it.

local S = require("qs.lib.std")
local qs = require("qs.globals")
local infer = require("qs.infer")
local progmod = require("qs.progmodule")

local libdir = os.getenv("QSB200_LIB_DIR") or "."
terralib.linklibrary(libdir .. "/libqsb200.so")
local B = terralib.includec("qsb200.h", {"-I", os.getenv("QSB200_INCLUDE_DIR") or "include"})

local C = terralib.includecstring [[
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
]]

-- S.assert-style failure: message, then abort (qs/lib/std.t:35-43)
local check = macro(function(rc)
	return quote
		if rc ~= 0 then
			C.printf("qsb200: %s\n", B.qsb_last_error())
			C.abort()
		end
	end
end)

-- ---------------------------------------------------------------------------------------------
-- Kernel constructors: same tables, same defaults as the reference.
local function kernelSpec(kind, doAdapt, stepSize, numSteps, scale, lexical)
	return { specs = { {kind=kind, doAdapt=doAdapt and 1 or 0, stepSize=stepSize, numSteps=numSteps, scale=scale, weight=1.0,
	                    lexicalScaleSharing=lexical and 1 or 0} } }
end

local function HMCKernel(params)            -- qs/hmc.t:57-66
	params = params or {}
	local doStepSizeAdapt = true
	if params.doStepSizeAdapt ~= nil then doStepSizeAdapt = params.doStepSizeAdapt end
	return kernelSpec(B.QSB_KERNEL_HMC, doStepSizeAdapt, params.stepSize or 1.0, params.numSteps or 1, 1.0)
end

local function DriftKernel(params)          -- qs/drift.t:63-73
	params = params or {}
	local doScaleAdapt = true
	if params.doScaleAdapt ~= nil then doScaleAdapt = params.doScaleAdapt end
	local lexicalScaleSharing = false           -- drift.t:71-73
	if params.lexicalScaleSharing ~= nil then lexicalScaleSharing = params.lexicalScaleSharing end
	return kernelSpec(B.QSB_KERNEL_DRIFT, doScaleAdapt, 1.0, 1, params.scale or 1.0, lexicalScaleSharing)
end

local function HARMKernel(params)           -- qs/harm.t:16-22
	params = params or {}
	local doScaleAdapt = true
	if params.doScaleAdapt ~= nil then doScaleAdapt = params.doScaleAdapt end
	return kernelSpec(B.QSB_KERNEL_HARM, doScaleAdapt, 1.0, 1, params.scale or 1.0)
end

local function TraceMHKernel(params)        -- qs/mcmc.t:134-193; fixed-structure programs: the non-structural half only
	params = params or {}
	assert(params.doStruct == false, "qs.b200.TraceMHKernel: pass {doStruct=false} (structural choices stay on qs.MCMC)")
	return kernelSpec(B.QSB_KERNEL_TRACEMH, false, 1.0, 1, 1.0)
end

local function MixtureKernel(kernels, weights)   -- qs/mcmc.t:199-203
	assert(#kernels == #weights, "MixtureKernel: number of kernels must equal number of weights")
	local specs = {}
	for i,k in ipairs(kernels) do
		assert(#k.specs == 1, "qs.b200.MixtureKernel: nested mixtures are not supported")
		assert(not k.annealSched, "qs.b200.MixtureKernel: anneal the mixture, not its parts")
		local s = {}
		for key,v in pairs(k.specs[1]) do s[key] = v end
		s.weight = weights[i]
		specs[#specs+1] = s
	end
	return { specs = specs }
end

-- qs/mcmc.t:297-319.  annealSched is any Terra function (iter: uint, numiters: uint) -> qs.float; it stays host
-- code: the run tabulates it once and hands the table to the device (qsb_sampler_set_temperatures).
local function AnnealingKernel(kernel, annealSched)
	assert(not kernel.annealSched, "qs.b200.AnnealingKernel: nested schedules are not supported")
	return { specs = kernel.specs, annealSched = annealSched }
end

-- ---------------------------------------------------------------------------------------------
-- Model descriptors.  The device path runs a closed set of fixed-structure families; a qs.program is tagged with
-- the descriptor of its family:  qs.b200.describe(program, {family=..., dim=..., ...}).  Pointers in the table are
-- Terra globals / constant arrays owned by the caller.
local descriptors = {}
local function describe(program, desc)
	progmod.assertIsProgram(program, "qs.b200.describe")
	descriptors[program] = desc
	return program
end

-- ---------------------------------------------------------------------------------------------
-- qs.b200.MCMC: the replacement of doMCMC (qs/mcmc.t:45-103).
--   numchains (default 1 = the reference's single chain, identical sample shape), seed, device: new optional parameters;
--   devices = {0, 1, ...}: the chains are block-partitioned over these GPUs of the box and driven from this one host thread
--   (qsb_group: the library owns one stream per device and the NCCL communicator of its diagnostics).
local function MCMC(kernel, params)
	params = params or {}
	local _numsamps = params.numsamps or 1000      -- mcmc.t:38
	local _burnin = params.burnin or 0             -- mcmc.t:39
	local _lag = params.lag or 1                   -- mcmc.t:40
	local _numchains = params.numchains or 1
	local _seed = params.seed or 0
	local _device = params.device or -1
	local _devices = params.devices                -- nil: a single device
	local _verbose = params.verbose and true or false
	local nspecs = #kernel.specs
	local ndev = _devices and #_devices or 0

	return function(program)
		progmod.assertIsProgram(program, "MCMC")
		local desc = descriptors[program]
		if not desc then
			error("qs.b200.MCMC: program has no device descriptor (structural or discrete programs stay on qs.MCMC)")
		end
		local SampleT = infer.SampleType(program)
		-- length of the return value (infer.t:47-49): an INDEP program returning sum/ret_div yields a scalar, everything else its vector
		local retdimL = (desc.family == B.QSB_FAM_INDEP and (desc.ret_mode or 1) == 0) and 1 or desc.dim
		-- How a sample's value is rebuilt from the doubles the library returns.  Sample(T) = {value: T, logprob, loglikelihood}
		-- (infer.t:13-18): T = double or double[n] is plain data and is copied; T = S.Vector(qs.float) (what the reference's
		-- vector-valued test programs return, testsuite.t:705-716) owns heap storage and is built element by element.
		-- Anything else is refused here, at compile time, instead of being overwritten with raw doubles.
		local ValueT
		for _,e in ipairs(SampleT.entries) do if e.field == "value" or e[1] == "value" then ValueT = e.type or e[2] end end
		local valueIsVector = (ValueT == S.Vector(qs.float))
		local valueIsPOD = (ValueT == double) or (ValueT:isarray() and ValueT.type == double)
		if not (valueIsVector or valueIsPOD) then
			error("qs.b200.MCMC: the program's return type must be double, double[n] or S.Vector(qs.float)")
		end

		return terra(samples: &S.Vector(SampleT))
			var d : B.qsb_model_desc
			C.memset(&d, 0, sizeof(B.qsb_model_desc))
			d.family = [desc.family]; d.dim = [desc.dim]; d.nobs = [desc.nobs or 0]
			d.ret_mode = [desc.ret_mode or 1]; d.ret_div = [desc.ret_div or 1.0]; d.prior_sigma = [desc.prior_sigma or 0.0]
			d.erp_kind = [desc.erp_kind or `nil]; d.erp_p0 = [desc.erp_p0 or `nil]; d.erp_p1 = [desc.erp_p1 or `nil]
			d.A = [desc.A or `nil]; d.X = [desc.X or `nil]; d.ybin = [desc.ybin or `nil]
			d.y = [desc.y or `nil]; d.sigma = [desc.sigma or `nil]; d.erp_lexid = [desc.erp_lexid or `nil]
			d.erp_fmu = [desc.erp_fmu or `nil]; d.erp_fsigma = [desc.erp_fsigma or `nil]
			d.erp_pref = [desc.erp_pref or `nil]; d.erp_pcoef = [desc.erp_pcoef or `nil]; d.erp_ptf = [desc.erp_ptf or `nil]
			d.erp_cond_lo = [desc.erp_cond_lo or `nil]; d.erp_cond_hi = [desc.erp_cond_hi or `nil]

			var specs : B.qsb_kernel_spec[ nspecs ]
			escape
				for i,s in ipairs(kernel.specs) do
					emit quote
						specs[ [i-1] ].kind = [s.kind]; specs[ [i-1] ].doAdapt = [s.doAdapt]
						specs[ [i-1] ].stepSize = [s.stepSize]; specs[ [i-1] ].numSteps = [s.numSteps]
						specs[ [i-1] ].scale = [s.scale]; specs[ [i-1] ].weight = [s.weight]
						specs[ [i-1] ].lexicalScaleSharing = [s.lexicalScaleSharing or 0]; specs[ [i-1] ].reserved = 0
					end
				end
			end
			var cfg : B.qsb_run_config
			cfg.seed = [_seed]; cfg.chain_begin = 0; cfg.numchains = [_numchains]; cfg.sample_chains = 0
			cfg.flags = 0; cfg.device = [_device]; cfg.reserved = 0; cfg.stream = nil

			var model : &B.qsb_model = nil
			var sampler : &B.qsb_sampler = nil
			var group : &B.qsb_group = nil
			var iters : uint = [_burnin] + [_numsamps] * [_lag]
			var temps = [S.Vector(double)].salloc():init()
			escape
				if kernel.annealSched then   -- AnnealingKernel:next calls annealSched(iter, numiters) (mcmc.t:308-311)
					local sched = kernel.annealSched
					emit quote for i=0,iters do temps:insert(sched(i, iters)) end end
				end
				if ndev > 0 then
					emit quote
						var devs : int32[ ndev ]
						escape for i,dv in ipairs(_devices) do emit quote devs[ [i-1] ] = [dv] end end end
						check(B.qsb_group_create(&d, specs, nspecs, &cfg, devs, ndev, &group))
						check(B.qsb_group_init(group))                                  -- TraceType.salloc():init(true), mcmc.t:55
						if temps:size() > 0 then
							for g=0,B.qsb_group_size(group) do
								check(B.qsb_sampler_set_temperatures(B.qsb_group_sampler(group, g), temps:get(0), iters))
							end
						end
						check(B.qsb_group_run(group, [_numsamps], [_burnin], [_lag]))   -- the for i=0,iters loop, mcmc.t:57-72
					end
				else
					emit quote
						check(B.qsb_model_create(&d, [_device], &model))
						check(B.qsb_sampler_create(model, specs, nspecs, &cfg, &sampler))
						check(B.qsb_sampler_init(sampler))
						if temps:size() > 0 then check(B.qsb_sampler_set_temperatures(sampler, temps:get(0), iters)) end
						check(B.qsb_sampler_run(sampler, [_numsamps], [_burnin], [_lag]))
					end
				end
			end

			-- Caller owns the vector (qs/infer.t:127) and the reference APPENDS to it (samples:insert(), mcmc.t:65): the kept
			-- samples arrive as rows of (retdim + 2) doubles {value.., logprob, loglikelihood} in a scratch buffer and every
			-- Sample is constructed in place, chain after chain.
			var n : uint64 = 0
			var retdim : int = [retdimL]
			escape
				if ndev > 0 then emit quote n = B.qsb_sampler_num_samples(B.qsb_group_sampler(group, 0)) end
				else emit quote n = B.qsb_sampler_num_samples(sampler) end end
			end
			var width = retdim + 2
			var rows = [&double](C.malloc(n * [_numchains] * width * sizeof(double)))
			escape
				if ndev > 0 then
					emit quote check(B.qsb_group_fetch_samples(group, 0, [_numchains], 0, n, [&opaque](rows), width * sizeof(double))) end
				else
					emit quote check(B.qsb_sampler_fetch_samples(sampler, 0, [_numchains], 0, n, [&opaque](rows), width * sizeof(double))) end
				end
			end
			for i=0, n * [_numchains] do
				var row = rows + i * width
				var s = samples:insert()
				escape
					if valueIsVector then
						emit quote
							s.value:init()
							for j=0,retdim do s.value:insert(row[j]) end
						end
					else
						emit quote C.memcpy(&s.value, row, retdim * sizeof(double)) end
					end
				end
				s.logprob = row[retdim]
				s.loglikelihood = row[retdim + 1]
			end
			C.free(rows)

			escape
				if _verbose then
					emit quote
						var st : B.qsb_stats
						escape
							if ndev > 0 then emit quote check(B.qsb_group_stats(group, &st)) end
							else emit quote check(B.qsb_sampler_stats(sampler, &st)) end end
						end
						-- same line formats as mcmc.t:81,88 (the counters are 64-bit here)
						C.printf("Acceptance Ratio: %lu/%lu (%g%%)\n", st.props_accepted, st.props_made,
							100.0*double(st.props_accepted)/double(st.props_made))
						C.printf("Time: %g\n", st.seconds)
					end
				end
			end
			if group ~= nil then check(B.qsb_group_destroy(group)) end
			if sampler ~= nil then check(B.qsb_sampler_destroy(sampler)) end
			if model ~= nil then check(B.qsb_model_destroy(model)) end
		end
	end
end

return
{
	exports =
	{
		b200 =
		{
			MCMC = MCMC,
			HMCKernel = HMCKernel,
			DriftKernel = DriftKernel,
			HARMKernel = HARMKernel,
			TraceMHKernel = TraceMHKernel,
			MixtureKernel = MixtureKernel,
			AnnealingKernel = AnnealingKernel,
			describe = describe,
			FAM_INDEP = B.QSB_FAM_INDEP, FAM_GAUSS_PREC = B.QSB_FAM_GAUSS_PREC, FAM_LOGISTIC = B.QSB_FAM_LOGISTIC,
			FAM_EIGHT_SCHOOLS = B.QSB_FAM_EIGHT_SCHOOLS, FAM_FUNNEL = B.QSB_FAM_FUNNEL
		}
	}
}
