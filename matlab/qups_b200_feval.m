function y = qups_b200_feval(op, C, varargin)
% QUPS_B200_FEVAL - drop-in for parallel.gpu.CUDAKernel.feval at the hot-path call sites of QUPS
%
% y = QUPS_B200_FEVAL('das', C, yg, Pi, Pr, Pv, Nv, apod, cinv, [cstride, astride], x, [fs, fmod])
% replaces (kern/das_spec.m:371-373)
%     y{f} = k.feval(yg, Pi, Pr, Pv, Nv, apod, cinv, [cstride, astride], x(:,:,:,f), [fs, fmod]);
% C is a struct of the values the reference writes with k.setConstantMemory (kern/das_spec.m:294-298):
%     C = struct('I1',Isz(1),'I2',Isz(2),'I3',Isz(3),'N',N,'M',M,'T',T,'S',S,'VS',VS,'DV',DV,'flag',flagnum);
%
% y = QUPS_B200_FEVAL('ws2', C, y_, w_, x_, t1_, t2_, uint64(dsizes), uint64(strides))
% replaces (kern/wsinterpd2.m:235); C = struct('T',T,'interp',flagnum,'omega',imag(omega)).
%
% x = QUPS_B200_FEVAL('greens', C, x, ps, as, pn, pv, kn)
% replaces (src/UltrasoundSystem.m:718); C = struct('n0',t(1),'t0x',wv.t0,'fs',fso,'fsr',wv.fs/fso, ...
%     'c0',c0,'R0',kwargs.R0,'E',E,'interp',flagnum).  The host-side scatterer windowing (sb, iblock,
%     :678-714) is not needed: the library windows internally.
%
% y = QUPS_B200_FEVAL('das', C, ..., [fs, fmod], rx_aux, tx_aux, lat) with C.ap_rx_kind / ap_tx_kind / ap_rx_p / ap_tx_p /
% ap_lat_dim set from a qups_b200_apod struct evaluates closed-form apodization inside the kernel
% (replaces the dense arrays of src/UltrasoundSystem.m:4892-5429); a = QUPS_B200_FEVAL('apod', C, a0, Pi, Pr, rx_aux,
% tx_aux, lat) returns the dense array itself; y = QUPS_B200_FEVAL('prep', C, y0, x, t0) fuses
% zeropad -> hilbert -> downmix -> cast (src/ChannelData.m:757-807, 935-966, 1153-1183) into one pass.
%
% All array arguments are gpuArrays of the class the reference already prepares at these sites; the result
% is a gpuArray with the size and class of the first array argument, like feval's return value.
%
% Requires the MEX gateway built from mex/qups_b200_mex.cu (see INTEGRATION.md).
    y = qups_b200_mex(op, C, varargin{:});
end
