function [q,qdot,status,iters] = rolloutBatch(scene,scheme,q0,qdot0,tau)
% rolloutBatch  GPU replacement of simLoop(scene) of driverRedMaxBDF1.m:57 / driverRedMaxBDF2.m:57 for B rollouts.
%   scene   redmax.Scene after scene.init()  (built by scenesRedMax.m unchanged)
%   scheme  1 = BDF1, 2 = SDIRK2 start + BDF2
%   q0,qdot0  nr x B initial states (column b = rollout b); tau: [] | nr x B | nr x nsteps x B
%   q,qdot  nr x nsteps x B == history(k).q / history(k).qdot of each rollout (Scene.m:136-137)
% See INTEGRATION.md.  Requires redmax_mex (matlab/redmax_mex.cpp) and libredmax_b200.so.
if nargin < 5, tau = []; end
h = redmax_mex('create',flattenScene(scene));
opts = struct('scheme',scheme,'nsteps',scene.nsteps,'h',scene.h);
[q,qdot,status,iters] = redmax_mex('rollout',h,opts,q0,qdot0,tau);
redmax_mex('destroy',h);
if scheme == 2
	[q,qdot,status,iters] = reparamRollouts(scene,opts,q0,qdot0,tau,q,qdot,status,iters);
end
end

function [q,qdot,status,iters] = reparamRollouts(scene,opts,q0,qdot0,tau,q,qdot,status,iters)
% jroot.reparam() (driverRedMaxBDF2.m:112) for the rollouts the library flagged with RMX_ST_CHART (bit 32): cut the rollout at
% the first step whose result leaves the well-conditioned range of a JointSpherical chart, let the joint objects themselves
% re-express that step and the BDF2 history (jroot.reparam() -> JointSpherical.reparam_, JointSpherical.m:63-103), rebuild the scene description
% with the new charts and resume from the next step.  Same loop as redmax_b200/scene.py _reparam_rollouts, one rollout at a
% time for clarity.  After a switch q(:,k,b) is expressed in the new chart, as history(k).q is in the reference.
sph = scene.joints(cellfun(@(j) isa(j,'redmax.JointSpherical'), scene.joints));
if isempty(sph), return; end
chart0 = cellfun(@(j) j.chart, sph);
z = zeros(3,1);
for b = find(bitand(status(:)',32))
	for i = 1 : length(sph), sph{i}.chart = chart0(i); end
	kb = 0; st = int32(0); it = zeros(2,1,'int32'); hist = []; % hist: {step (0-based, -1 = initial state), q1, qdot1}
	qb = q(:,:,b); qdb = qdot(:,:,b); taub = [];
	if ~isempty(tau) % constant torque: nr x B ; per step: nr x nsteps x B
		if numel(tau) == size(q,1)*size(q,3), taub = tau(:,b); else, taub = tau(:,:,b); end % same rule as the gateway
	end
	while kb < opts.nsteps
		% first step (0-based) at or after kb with |det T| <= 0.5 for some spherical joint
		k1 = opts.nsteps - 1; sw = false;
		for k = kb : opts.nsteps-1
			for i = 1 : length(sph)
				[~,~,~,~,~,detT] = redmax.JointSpherical.getEuler(sph{i}.chart,qb(sph{i}.idxR,k+1),z);
				sw = sw || abs(detT) <= 0.5;
			end
			if sw, k1 = k; break; end
		end
		% the piece [kb, k1] on its own: its status and iteration counts, without the discarded tail
		o = opts; o.nsteps = k1 + 1;
		[~,~,s1,i1] = resumeWith(scene,o,kb,q0(:,b),qdot0(:,b),sliceTau(taub,o.nsteps),qb(:,1:k1+1),qdb(:,1:k1+1),hist);
		st = bitor(st,bitand(s1,int32(bitcmp(uint32(32))))); it = it + i1;
		if ~sw, break; end
		% the joint objects re-express step k1 and the history (joint.q1 = step k1-1, or the initial state)
		if ~isempty(hist) && hist{1} == k1-1, hq = hist{2}; hqd = hist{3};
		elseif k1 == 0, hq = q0(:,b); hqd = qdot0(:,b);
		else, hq = qb(:,k1); hqd = qdb(:,k1); end
		for i = 1 : length(sph)
			j = sph{i}; r = j.idxR;
			j.q = qb(r,k1+1); j.qdot = qdb(r,k1+1); j.q1 = hq(r); j.qdot1 = hqd(r); j.chart1 = j.chart;
		end
		scene.joints{1}.reparam(); % Joint.m:372: reparam_ is protected; the walk is a no-op for every other joint type
		for i = 1 : length(sph)
			j = sph{i}; r = j.idxR;
			qb(r,k1+1) = j.q; qdb(r,k1+1) = j.qdot; hq(r) = j.q1; hqd(r) = j.qdot1;
		end
		hist = {k1-1,hq,hqd}; kb = k1 + 1;
		if kb < opts.nsteps
			[qb,qdb] = resumeWith(scene,opts,kb,q0(:,b),qdot0(:,b),taub,qb,qdb,hist);
		end
	end
	q(:,:,b) = qb; qdot(:,:,b) = qdb; status(b) = st; iters(:,b) = it;
end
for i = 1 : length(sph), sph{i}.chart = chart0(i); end
end

function [qb,qdb,st,it] = resumeWith(scene,opts,kb,q0,qdot0,tau,qb,qdb,hist)
% redmax_mex('resume') under the charts the joint objects hold now, with the re-expressed BDF2 history in place of the stored step
keep = [];
if ~isempty(hist)
	if hist{1} < 0, q0 = hist{2}; qdot0 = hist{3};
	else, keep = {qb(:,hist{1}+1),qdb(:,hist{1}+1)}; qb(:,hist{1}+1) = hist{2}; qdb(:,hist{1}+1) = hist{3}; end
end
h = redmax_mex('create',flattenScene(scene));
[qb,qdb,st,it] = redmax_mex('resume',h,opts,double(kb),q0,qdot0,tau,qb,qdb);
redmax_mex('destroy',h);
if ~isempty(keep), qb(:,hist{1}+1) = keep{1}; qdb(:,hist{1}+1) = keep{2}; end
end

function t = sliceTau(tau,ns)
t = tau; if size(tau,2) > 1, t = tau(:,1:ns); end
end

function d = flattenScene(scene)
% Flattens the +redmax object graph into the arrays of rmx_scene_desc (include/redmax_b200.h).
n = length(scene.joints);
d.parent = zeros(1,n); d.jtype = zeros(1,n);
d.E0_pj = zeros(4,4,n); d.E0_ji = zeros(4,4,n); d.axis = zeros(3,n); d.I_i = zeros(6,n); d.sides = zeros(3,n);
d.axis2 = repmat([0;1;0],1,n);
d.stiffness = zeros(1,n); d.damping = zeros(1,n); d.qRest = zeros(6,n); % RMX_MAX_JOINT_DOF x n
d.chart = zeros(1,n); % double like every other index array (the gateway also accepts int32)
d.qLimL = zeros(1,n); d.qLimU = zeros(1,n); d.qLimK = zeros(1,n); d.qLimD = zeros(1,n);
for i = 1 : n
	j = scene.joints{i};
	if isempty(j.parent)
		d.parent(i) = -1;
	else
		d.parent(i) = find(cellfun(@(x) x == j.parent, scene.joints)) - 1;
	end
	d.qRest(1:j.ndof,i) = j.qRest;
	if isa(j,'redmax.JointRevolute')
		d.jtype(i) = 1; d.axis(:,i) = j.axis;
	elseif isa(j,'redmax.JointFixed')
		d.jtype(i) = 0;
	elseif isa(j,'redmax.JointPrismatic')
		d.jtype(i) = 2; d.axis(:,i) = j.axis;
	elseif isa(j,'redmax.JointPlanar')
		d.jtype(i) = 3; d.axis(:,i) = j.plane(:,1); d.axis2(:,i) = j.plane(:,2);
	elseif isa(j,'redmax.JointTranslational')
		d.jtype(i) = 4;
	elseif isa(j,'redmax.JointFree2D')
		d.jtype(i) = 5;
	elseif isa(j,'redmax.JointUniversal')
		d.jtype(i) = 6;
	elseif isa(j,'redmax.JointSpherical')
		% status bit 32 marks rollouts that need jroot.reparam(): re-express that step, rebuild d with the new charts, 'resume'
		d.jtype(i) = 7; d.chart(i) = j.chart;
	elseif isa(j,'redmax.JointFree3D')
		d.jtype(i) = 8; d.chart(i) = j.joint2.chart;
	else
		error('joint type %s is not on the GPU hot path',class(j));
	end
	d.E0_pj(:,:,i) = j.E0_pj; d.E0_ji(:,:,i) = j.body.E0_ji; d.I_i(:,i) = j.body.I_i; d.sides(:,i) = j.body.sides;
	d.stiffness(i) = j.stiffness; d.damping(i) = j.damping;
	d.qLimL(i) = j.qLimL; d.qLimU(i) = j.qLimU; d.qLimK(i) = j.qLimK; d.qLimD(i) = j.qLimD;
end
d.grav = scene.grav;
gb = []; gE = zeros(4,4,0); kn = []; kt = []; kd = []; mu = [];
pb1 = []; pb2 = []; px1 = zeros(3,0); px2 = zeros(3,0); pks = []; pkd = []; pkind = []; pL = [];
cn = []; cb = zeros(4,0); cx = zeros(3,4,0); cks = []; ckd = []; cL = []; % RMX_MAX_CABLE_POINTS = 4
for i = 1 : length(scene.forces)
	f = scene.forces{i};
	if isa(f,'redmax.ForceGroundCuboid')
		gb(end+1) = find(cellfun(@(x) x == f.cuboid, scene.bodies)) - 1; %#ok<AGROW>
		gE(:,:,end+1) = f.E; kn(end+1) = f.kn; kt(end+1) = f.kt; kd(end+1) = f.kd; mu(end+1) = f.mu; %#ok<AGROW>
	elseif isa(f,'redmax.ForcePointPoint') || isa(f,'redmax.ForceSpringDamper')
		pb1(end+1) = bodyIndex(scene,f.body1); pb2(end+1) = bodyIndex(scene,f.body2); %#ok<AGROW>
		px1(:,end+1) = f.x_1; px2(:,end+1) = f.x_2; pks(end+1) = f.stiffness; pkd(end+1) = f.damping; %#ok<AGROW>
		if isa(f,'redmax.ForceSpringDamper')
			pkind(end+1) = 1; pL(end+1) = f.L; %#ok<AGROW> % L was set by scene.init() (ForceSpringDamper.m:38-62)
		else
			pkind(end+1) = 0; pL(end+1) = 0; %#ok<AGROW>
		end
	elseif isa(f,'redmax.ForceCable')
		np = length(f.bodies);
		cn(end+1) = np; cb(:,end+1) = -1; cx(:,:,end+1) = 0; %#ok<AGROW>
		for k = 1 : np
			cb(k,end) = bodyIndex(scene,f.bodies{k}); cx(:,k,end) = f.xls{k};
		end
		cks(end+1) = f.stiffness; ckd(end+1) = f.damping; cL(end+1) = f.L; %#ok<AGROW>
	elseif ~isa(f,'redmax.ForceNull')
		error('only ForceGroundCuboid, ForcePointPoint, ForceSpringDamper and ForceCable are on the GPU hot path');
	end
end
d.pf_body1 = pb1; d.pf_body2 = pb2; d.pf_x1 = px1; d.pf_x2 = px2; d.pf_ks = pks; d.pf_kd = pkd; d.pf_kind = pkind; d.pf_L = pL;
d.cable_npts = cn; d.cable_body = cb; d.cable_x = cx; d.cable_ks = cks; d.cable_kd = ckd; d.cable_L = cL;
d.ground_body = gb; d.ground_E = gE; d.ground_kn = kn; d.ground_kt = kt; d.ground_kd = kd; d.ground_mu = mu;
end

function i = bodyIndex(scene,body)
% 0-based body index, -1 for the world (empty body)
if isempty(body)
	i = -1;
else
	i = find(cellfun(@(x) x == body, scene.bodies)) - 1;
end
end
